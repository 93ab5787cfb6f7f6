"""Multi-GPU plumbing: the planning path shards across INDEPENDENT planning problems only
 -- one process per GPU, contiguous block partition of the environments,
no data-path collective inside a plan.  NCCL (via torch.distributed) is used for exactly two
things: broadcasting the TOLD parameters after an update, and gathering the planned actions.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block partition: rank r plans problems [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights(model, src=0):
    """After agent.update() on the trainer rank: one flat broadcast of every TOLD parameter AND floating-point buffer
    (about 6 MB fp32), then each rank re-packs locally (Planner.sync_weights).  Buffers matter: the GRU agent's prior MLP
    carries BatchNorm running statistics (`_dynamics.prior_mlp.1.running_mean / running_var`) that update() moves in
    train mode and the planner folds into its packed weights."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = list(model.parameters()) + [b for b in model.buffers() if b.is_floating_point()]
    with torch.no_grad():
        flat = torch.cat([p.reshape(-1).float() for p in tensors])
        dist.broadcast(flat, src=src)
        off = 0
        for p in tensors:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))     # on the tensor itself (not .data): bumps Tensor._version -> planner re-packs
            off += n


def gather_actions(local_actions, total):
    """All-gather of the per-rank [B_r, A] action blocks into [total, A] (64 KB at B=4096, A=4:
    latency-bound).  Ranks may hold unequal shards (padded to the largest)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_actions
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = local_actions.new_zeros(cap, local_actions.shape[1])
    pad[: local_actions.shape[0]] = local_actions
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


class ShardedPlanner:
    """Plans `total` independent problems across the ranks of the default process group."""

    def __init__(self, planner, total):
        self.planner = planner
        self.total = total
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.lo, self.hi = shard_range(total, self.rank, self.world)
        # every rank would otherwise draw the SAME Philox stream for its local problem i: fold the global offset of this
        # shard into the seed (a 64-bit odd multiplier keeps distinct shards on distinct keys)
        self.base_seed = int(getattr(planner, "seed", 0))
        if hasattr(planner, "seed") and self.world > 1:
            planner.seed = (self.base_seed + 0x9E3779B97F4A7C15 * (self.lo + 1)) & 0xFFFFFFFFFFFFFFFF

    def plan(self, obs_all_or_local, gather=True, **kw):
        obs = obs_all_or_local
        if obs.shape[0] == self.total and self.world > 1:
            obs = obs[self.lo:self.hi]
        res = self.planner.plan(obs, **kw)
        if gather:
            res["actions_all"] = gather_actions(res["action"], self.total)
        return res

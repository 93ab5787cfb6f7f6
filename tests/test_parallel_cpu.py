"""Host-side multi-GPU logic on CPU: gloo backend, world_size 2 (problems shard, nothing else)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tdmpc_b200.parallel import ShardedPlanner, broadcast_weights, gather_actions, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 2, 7, 4096, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakePlanner:
    """Stands in for the CUDA planner: the 'action' of problem i is a function of its observation only."""
    seed = 7

    def plan(self, obs, **kw):
        return {"action": torch.stack([obs.sum(1), obs[:, 0] * 2], 1), "metrics": torch.zeros(obs.shape[0], 2)}


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        obs_all = torch.arange(total * 3, dtype=torch.float32).view(total, 3)
        sp = ShardedPlanner(_FakePlanner(), total)
        assert (sp.lo, sp.hi) == shard_range(total, rank, world)
        res = sp.plan(obs_all)
        ref = _FakePlanner().plan(obs_all)["action"]
        assert res["action"].shape[0] == sp.hi - sp.lo
        assert torch.equal(res["actions_all"], ref)                      # same per-problem result as one process
        assert torch.equal(gather_actions(res["action"], total), ref)
        # weight broadcast: rank 0's parameters win, in place (so Tensor._version changes -> re-pack)
        m = torch.nn.Linear(4, 3)
        with torch.no_grad():
            m.weight.fill_(float(rank + 1)); m.bias.fill_(float(10 * (rank + 1)))
        v0 = m.weight._version
        broadcast_weights(m, src=0)
        assert torch.all(m.weight == 1.0) and torch.all(m.bias == 10.0) and m.weight._version > v0
        # buffers travel too: the GRU agent's prior MLP folds BatchNorm running statistics into the packed weights
        bn = torch.nn.BatchNorm1d(5)
        with torch.no_grad():
            bn.running_mean.fill_(float(rank + 2)); bn.running_var.fill_(float(3 * rank + 4))
        vb = bn.running_var._version
        broadcast_weights(bn, src=0)
        assert torch.all(bn.running_mean == 2.0) and torch.all(bn.running_var == 4.0) and bn.running_var._version > vb
        assert bn.num_batches_tracked.dtype == torch.int64                      # integer buffers are left alone
        # each rank plans its shard with its own Philox key (same key => local problem i repeats rank 0's noise)
        seeds = [None] * world
        dist.all_gather_object(seeds, sp.planner.seed)
        assert len(set(seeds)) == world and sp.base_seed == 7
    finally:
        dist.destroy_process_group()


def test_sharded_plan_and_collectives_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 7), nprocs=2, join=True)       # 7 problems over 2 ranks: unequal shards

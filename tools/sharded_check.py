"""NCCL check of the multi-GPU plumbing on real GPUs (the gloo twin runs on CPU in tests/test_parallel_cpu.py):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/sharded_check.py
7 planning problems over the ranks (unequal shards): every rank plans its block with the CUDA planner, the all-gathered
actions must equal, bit for bit, the same 7 problems planned in one batch on one GPU and the CPU oracle to 1e-5; then the
TOLD weights of rank 0 are broadcast and the planners re-pack."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import gpu_util as gu, synth  # noqa: E402
from tdmpc_b200 import TOLD  # noqa: E402
from tdmpc_b200.parallel import ShardedPlanner, broadcast_weights, shard_range  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = 7
cfg = synth.make_cfg("tiny")
sd = synth.make_told_weights(cfg, 0)
obs = np.stack([synth.make_obs(cfg, 20 + i) for i in range(total)])
nzs = [synth.make_noise(cfg, 30 + i) for i in range(total)]
for prec in ("fp32", "bf16"):
    pl = gu.make_planner(cfg, sd, prec)
    kw = dict(**gu.plan_kwargs(cfg), eval_mode=True, t0=True)
    whole = pl.plan(torch.from_numpy(obs).cuda(), noise=gu.batch_noise(nzs), **kw)["action"].clone()
    sp = ShardedPlanner(pl, total)
    lo, hi = shard_range(total, rank, world)
    assert (sp.lo, sp.hi) == (lo, hi)
    res = sp.plan(torch.from_numpy(obs[lo:hi]).cuda(), noise=gu.batch_noise(nzs[lo:hi]), **kw)
    torch.cuda.synchronize()
    assert res["actions_all"].shape == whole.shape and torch.equal(res["actions_all"], whole), (prec, rank)
    if prec == "fp32":
        for i in range(lo, hi):
            a_o = gu.oracle_plan(sd, cfg, obs[i], nzs[i], t0=True, prev_mean=None)[0]
            assert torch.allclose(res["actions_all"][i].cpu(), a_o, rtol=1e-5, atol=1e-6), (i, res["actions_all"][i], a_o)
    # weight broadcast: rank r starts from different parameters; after the broadcast every rank plans with rank 0's
    m = TOLD(cfg).cuda()
    m.load_state_dict(synth.make_told_weights(cfg, rank))
    broadcast_weights(m, src=0)
    pl.sync_weights(m)
    again = pl.plan(torch.from_numpy(obs).cuda(), noise=gu.batch_noise(nzs), **kw)["action"]
    assert torch.equal(again, whole), (prec, rank, "broadcast weights differ from rank 0's")
    if rank == 0:
        print(f"sharded_check[{prec}] world={world}: all-gathered actions == single-batch plan (bit-exact), weight broadcast ok")
dist.barrier()
dist.destroy_process_group()

"""Batched iCEM planning sweep (BASELINE.json configs[3] / configs[4] dims) through Planner.plan:
    python tools/icem_batched.py --batches 256 1024 8192 [--mlp]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/icem_batched.py --batches 8192
Under torchrun every rank plans `batch` environments of its own on its own GPU (weak scaling over independent planning
problems, no data-path collective); the time is the max over ranks, the rate the whole job's."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, nargs="+", default=[256])
ap.add_argument("--precision", default="bf16")
ap.add_argument("--mlp", action="store_true")
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
for B in a.batches:
    if world > 1:
        dist.barrier()
    r = bench.icem_batched(dev, a.precision, B, not a.mlp, reps=2)
    if world > 1:
        t = torch.tensor([r["ms_per_batched_plan"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        scale = B * world / ms * 1e3 / r["plans_per_sec"]
        r.update(ms_per_batched_plan=ms, plans_per_sec=B * world / ms * 1e3, sample_steps_per_sec=r["sample_steps_per_sec"] * scale,
                 tflops=r["tflops"] * scale, n_gpus=world, envs_total=B * world)
    if rank == 0:
        print(json.dumps(r), flush=True)
if world > 1:
    dist.destroy_process_group()

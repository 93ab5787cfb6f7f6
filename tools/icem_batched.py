"""Batched iCEM planning sweep (BASELINE.json configs[3] / configs[4] dims) through Planner.plan on one GPU:
    python tools/icem_batched.py --batches 256 1024 8192 [--mlp]"""
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
for B in a.batches:
    print(json.dumps(bench.icem_batched(torch.device("cuda:0"), a.precision, B, not a.mlp, reps=2)), flush=True)

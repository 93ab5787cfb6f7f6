"""Per-job timeline of one 128-trajectory tile inside rollout_tc_kernel (clock64 stamps of the MMA
warp and the two epilogue groups).  python tools/trace_rollout.py [--batch 512] [--workload cfg3]"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth
from tdmpc_b200 import Planner, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="cfg3"); ap.add_argument("--batch", type=int, default=512)
a = ap.parse_args()
cfg = synth.make_cfg(a.workload); sd = synth.make_told_weights(cfg, 0); H, N, P, T = synth.plan_dims(cfg)
pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, precision="bf16")
pl.sync_weights({k: v.cuda() for k, v in sd.items()})
lib = _lib.load(); lib.tdmpc_debug_set_trace.argtypes = [C.c_void_p]; lib.tdmpc_debug_set_trace.restype = C.c_int
n = lib.tdmpc_debug_set_trace(None)
buf = torch.zeros(n, dtype=torch.int64, device="cuda")
z0 = pl.encode(torch.from_numpy(synth.make_obs(cfg, 1, batch=a.batch)).cuda())
acts = torch.rand(a.batch, T, H, cfg.action_dim, device="cuda") * 2 - 1
pl.estimate_value(z0, acts, discount=cfg.discount, min_std=cfg.min_std, seed=1, call=1)
lib.tdmpc_debug_set_trace(C.c_void_p(buf.data_ptr()))
pl.estimate_value(z0, acts, discount=cfg.discount, min_std=cfg.min_std, seed=1, call=1)
torch.cuda.synchronize(); lib.tdmpc_debug_set_trace(None)
tr = buf.cpu().numpy().reshape(3, -1, 2)
names = {0: "L1", 1: "L2", 2: "L3", 3: "QL1", 4: "QL2"}; chains = ["dyn", "rew", "pi", "q1", "q2"]
# rebuild the job list exactly as build_jobs(MODE_VALUE, H) does
jobs = []
for t in range(H):
    jobs += [(0, 0, c, t) for c in range(4)] + [(1, 0, c, t) for c in range(4)] + [(0, 1, c, t) for c in range(4)] + [(2, 0, 0, t)] + [(1, 1, c, t) for c in range(4)]
jobs += [(0, 2, c, 0) for c in range(4)] + [(1, 2, c, 0) for c in range(4)] + [(2, 2, 0, 0), (3, 3, 0, 0), (4, 3, 0, 0), (3, 4, 0, 0), (4, 4, 0, 0)]
t0 = tr[0, 0, 0]
print(f"{'job':>4} {'what':14} | {'mma start':>9} {'mma end':>9} {'dur':>6} | {'epi0 start':>10} {'end':>8} {'dur':>6} | {'epi1 start':>10} {'end':>8} {'dur':>6}")
for j, (ty, chn, c, t) in enumerate(jobs):
    row = [f"{j:4d} {names[ty] + ' ' + chains[chn] + ' c' + str(c) + ' t' + str(t):14}"]
    for r in range(3):
        s, e = tr[r, j]
        row.append(f"{(s - t0) if s else 0:>10} {(e - t0) if e else 0:>8} {(e - s) if s and e else 0:>6}")
    print(" | ".join(row))
print("tile total cycles:", tr[:, :len(jobs), 1].max() - t0)

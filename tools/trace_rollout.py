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
tr = buf.cpu().numpy().reshape(3, -1, 3)
names = {0: "L1", 1: "L2", 2: "L3"}; chains = ["dyn", "rew", "pi", "q1", "q2"]
epis = ["H1", "XR", "XS", "DOT", "Z", "PI", "STAT", "QN", "QDOT"]
# rebuild the job list exactly as build_jobs(MODE_VALUE, H, xr=True) does (rollout_tc.cu)
jobs = []
def l1(ch, t, e): jobs.extend((0, e, ch, c, t) for c in range(4))
def l2(ch, t, e): jobs.extend((1, e, ch, c, t) for c in range(4))
if os.environ.get("TDMPC_TC_XR", "1") != "0":
    l1(0, 0, 0)
    for t in range(H):
        last = t == H - 1
        for c in range(4):
            jobs += [(1, 2, 0, c, t), (0, 1, 1, c, t)]
        jobs.append((2, 4, 0, 0, t))
        nxt = 2 if last else 0
        tn = 0 if last else t + 1
        jobs.append((1, 3, 1, 0, t))
        for c in (1, 2, 3):
            jobs += [(1, 3, 1, c, t), (0, 0, nxt, c - 1, tn)]
        jobs.append((0, 0, nxt, 3, tn))
else:
    for t in range(H):
        l1(0, t, 0); l2(0, t, 2); l1(1, t, 0); jobs.append((2, 4, 0, 0, t)); l2(1, t, 3)
    l1(2, 0, 0)
l2(2, 0, 2); jobs.append((2, 5, 2, 0, 0))
for qh in (3, 4):
    l1(qh, 0, 6); l1(qh, 0, 7); l2(qh, 0, 6); l2(qh, 0, 8)
t0 = tr[0, 0, 0]
print(f"{'job':>4} {'what':20} | {'mma start':>9} {'wait':>6} {'issue':>6} | {'epi0 start':>10} {'wait':>6} {'work':>6} | {'epi1 start':>10} {'wait':>6} {'work':>6}")
for j, (ty, ep, chn, c, t) in enumerate(jobs):
    row = [f"{j:4d} {names[ty] + ' ' + chains[chn] + ' c' + str(c) + ' t' + str(t) + ' ' + epis[ep]:20}"]
    for r in range(3):
        s, w, e = tr[r, j]
        row.append(f"{(s - t0) if s else 0:>10} {(w - s) if s and w else 0:>6} {(e - w) if w and e else 0:>6}")
    print(" | ".join(row))
print("tile total cycles:", tr[:, :len(jobs), 2].max() - t0)

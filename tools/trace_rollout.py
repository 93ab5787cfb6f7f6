"""Per-job timeline of one 128-trajectory tile inside rollout_tc_kernel (clock64 stamps of the MMA
warp and the two epilogue groups).  python tools/trace_rollout.py [--batch 512] [--workload cfg3]"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth
from tdmpc_b200 import Planner, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="cfg3"); ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--lib", default=None, help="variant build (tools/build_variant.sh), e.g. one compiled with -DTDMPC_TRACE_WAITS")
a = ap.parse_args()
if a.lib: _lib.LIB_PATH = os.path.abspath(a.lib)
cfg = synth.make_cfg(a.workload)
H, N, P, T = synth.plan_dims(cfg)
drnn = a.workload in ("cfg5", "tiny_drnn")
sd = synth.make_drnn_weights(cfg, 0) if drnn else synth.make_told_weights(cfg, 0)
pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, precision="bf16", enc_layernorm=drnn,
             dynamics="drnn" if drnn else "mlp", hidden_dim=cfg.hidden_dim if drnn else 0)
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
raw = buf.cpu().numpy()
MJ = (n - 1280) // 6
has_ext = False      # builds with the finer stamps (git history: eafe16c, -DTDMPC_TRACE_WAITS) append [2][MJ][2] + [MJ][3] words
blk = raw[-1024:].reshape(256, 4); raw = raw[:-1024]
ends = raw[-256:-128]; ns = raw[-128:][ends > 0]; ends = ends[ends > 0]
tr = raw[:-256].reshape(3, -1, 2)
# the job list the kernel walked (word layout: tc_model.cuh)
lib.tdmpc_debug_jobs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]; lib.tdmpc_debug_jobs.restype = C.c_int
words = (C.c_uint32 * 1024)()
nj = lib.tdmpc_debug_jobs(pl._h, 0, H, words, 1024)
names = ["L1", "L2", "L3", "QL1", "QL2", "G", "GH", "GI"]; chains = ["dyn", "rew", "pi", "q1", "q2"]
jobs = [(w & 7, (w >> 3) & 7, (w >> 6) & 3, (w >> 8) & 31) for w in words[:nj]]
t0 = tr[0, 0, 0]
print(f"{'job':>4} {'what':14} | {'mma start':>9} {'mma end':>9} {'dur':>6} | {'epi0 start':>10} {'end':>8} {'dur':>6} | {'epi1 start':>10} {'end':>8} {'dur':>6} ")
for j, (ty, chn, c, t) in enumerate(jobs):
    row = [f"{j:4d} {names[ty] + ' ' + chains[chn] + ' c' + str(c) + ' t' + str(t):14}"]
    for r in range(3):
        s_, e_ = tr[r, j]
        row.append(f"{(s_ - t0) if s_ else 0:>10} {(e_ - t0) if e_ else 0:>8} {(e_ - s_) if s_ and e_ else 0:>6}")
    print(" | ".join(row))
print("tile total cycles:", tr[:, :len(jobs), 1].max() - t0)
print("tile periods of block 0 (cycles):", " ".join(str(int(x)) for x in np.diff(ends)))
if len(ends) > 1: print(f"SM clock from globaltimer: {(ends[-1] - ends[0]) / max(1, ns[-1] - ns[0]) * 1e3:.0f} MHz over {len(ends) - 1} tiles")
used = blk[blk[:, 2] > 0]
if len(used):
    per = (used[:, 1] - used[:, 0]) / used[:, 2]
    print(f"per-block cycles per tile over {len(used)} blocks: min {per.min():.0f} median {np.median(per):.0f} max {per.max():.0f}; "
          f"block cycles min {(used[:, 1] - used[:, 0]).min()} max {(used[:, 1] - used[:, 0]).max()}; tiles per block {used[:, 2].min()}..{used[:, 2].max()}")
    slow = np.argsort(-per)[:8]
    print("slowest blocks (block, smid, cycles/tile):", [(int(i), int(used[i, 3]), int(per[i])) for i in slow])

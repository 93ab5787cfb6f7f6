"""Tiny driver for profiling: a few estimate_value launches (the dominant kernel) and one full plan."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth  # noqa: E402
from tdmpc_b200 import Planner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--plan", action="store_true")
a = ap.parse_args()
cfg = synth.make_cfg(a.workload)
sd = synth.make_told_weights(cfg, 0)
H, N, P, T = synth.plan_dims(cfg)
pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, precision=a.precision)
pl.sync_weights({k: v.cuda() for k, v in sd.items()})
obs = torch.from_numpy(synth.make_obs(cfg, 1, batch=a.batch)).cuda()
z0 = pl.encode(obs)
acts = torch.rand(a.batch, T, H, cfg.action_dim, device="cuda") * 2 - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps):
    e0.record()
    pl.estimate_value(z0, acts, discount=cfg.discount, min_std=cfg.min_std, seed=1, call=1)
    e1.record()
    torch.cuda.synchronize()
    print(f"estimate_value B={a.batch} T={T}: {e0.elapsed_time(e1):.3f} ms")
if a.plan:
    kw = dict(horizon=H, num_samples=N, num_pi_trajs=P, num_elites=cfg.num_elites, iterations=cfg.iterations,
              discount=cfg.discount, temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std,
              std_floor=cfg.min_std, eval_mode=True)
    pl.plan(obs, t0=True, **kw)
    torch.cuda.synchronize()
    print("plan done, launches", pl.launch_count)

"""Batched iCEM-DRNN planning throughput (BASELINE.json configs[4] dims) through Planner.plan: plans/s and the kernel mix."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth
from tdmpc_b200 import Planner
from tdmpc_b200.colored import synthesis_matrix
import numpy as np
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); ap.add_argument("--precision", default="bf16"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = synth.make_cfg("cfg5"); sd = synth.make_drnn_weights(cfg, 0)
pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, enc_layernorm=True, precision=a.precision,
             dynamics="drnn", hidden_dim=cfg.hidden_dim)
pl.sync_weights({k: v.cuda() for k, v in sd.items()})
B, H, K, I, A = a.batch, cfg.horizon, cfg.num_elites, cfg.iterations, cfg.action_dim
sizes, n = [], cfg.num_samples
for i in range(I):
    if i > 0: n = max(2 * K, int(n / cfg.factor_decrease_num))
    sizes.append((n, int(cfg.fraction_elites_reused * K), int(0.5 * n)))
m = np.zeros((3, H, H), dtype=np.float32); m[0] = m[1] = m[2] = synthesis_matrix(cfg.noise_beta, H)
color = {"len": H, "mats": torch.from_numpy(m).cuda(), "n_white": [0] * I, "n_color1": [s[0] for s in sizes]}
obs = torch.from_numpy(synth.make_obs(cfg, 1, batch=B)).cuda()
hid = torch.zeros(B, cfg.hidden_dim, device="cuda"); hout = torch.empty_like(hid)
ebuf = torch.zeros(B, K, H, A, device="cuda"); prev = torch.zeros(B, H, A, device="cuda")
kw = dict(horizon=H, num_samples=cfg.num_samples, num_pi_trajs=sizes[0][2], num_elites=K, iterations=I, discount=cfg.discount,
          temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std, std_floor=cfg.min_std, init_std=0.5, eval_mode=True,
          n_samples=[s[0] for s in sizes], n_reuse=[s[1] for s in sizes], n_pi=[s[2] for s in sizes], elite_actions=ebuf, elite_horizon=H,
          warm_last=True, best_a_last=True, shift_elites=True, color=color, hidden=hid, hidden_out=hout, prev_mean=prev)
pl.plan(obs, t0=False, has_prev=True, **kw); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(a.reps):
    e0.record(); pl.plan(obs, t0=False, has_prev=True, **kw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"iCEM-DRNN {a.precision} B={B}: {ms:.2f} ms per batched plan -> {B / ms * 1e3:.1f} plans/s, rows per iteration {[sum(s) for s in sizes]}")

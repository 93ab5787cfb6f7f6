"""Which part of the value carries the bf16-mode error?  Ablations on one config / seed (uniform action rows):
full model, Q heads zeroed (value = discounted rewards), reward head zeroed (value = gamma^H min Q), horizons 1..H."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg1"); ap.add_argument("--seed", type=int, default=0); ap.add_argument("--lib", default=None)
a = ap.parse_args()
from tdmpc_b200 import _lib
if a.lib: _lib.LIB_PATH = os.path.abspath(a.lib)
from oracle import tdmpc_oracle as O
from tests import synth, gpu_util as gu
cfg = synth.make_cfg(a.workload); sd0 = synth.make_told_weights(cfg, a.seed)
obs = synth.make_obs(cfg, a.seed * 100)
z0 = O.told_h(sd0, torch.from_numpy(obs).unsqueeze(0))
rs = np.random.RandomState(a.seed + 7)
Hfull, N, P, T = synth.plan_dims(cfg)
acts_full = torch.from_numpy(rs.uniform(-1, 1, (Hfull, T, cfg.action_dim)).astype(np.float32))
eps = torch.from_numpy(rs.standard_normal((T, cfg.action_dim)).astype(np.float32))
def zero(sd, keys):
    sd = {k: v.clone() for k, v in sd.items()}
    for k in keys: sd[k] = torch.zeros_like(sd[k])
    return sd
variants = {"full": sd0, "no-Q (rewards only)": zero(sd0, ["_Q1.6.weight", "_Q1.6.bias", "_Q2.6.weight", "_Q2.6.bias"]),
            "no-reward (Q only)": zero(sd0, ["_reward.4.weight", "_reward.4.bias"])}
for name, sd in variants.items():
    pl = gu.make_planner(cfg, sd, "bf16")
    for H in sorted({1, 2, Hfull}):
        acts = acts_full[:H]
        v_o, _ = O.estimate_value(sd, z0.repeat(T, 1), acts, H, cfg.discount, cfg.min_std, eps, True)
        v_f, _ = O.estimate_value(sd, z0.repeat(T, 1), acts, H, cfg.discount, cfg.min_std, eps, False)
        v_g, _ = pl.estimate_value(z0.cuda(), acts.permute(1, 0, 2).unsqueeze(0).contiguous(), discount=cfg.discount, min_std=cfg.min_std, noise_term=eps.unsqueeze(0))
        s = v_o.abs().max()
        err = (v_g[0].cpu() - v_o.squeeze(1)).abs() / s
        qn = (v_f - v_o).abs().squeeze(1) / s
        print(f"{a.workload} s{a.seed} {name:22s} H={H}: median {err.median().item():.2e} p75 {err.quantile(0.75).item():.2e} p90 {err.quantile(0.9).item():.2e} "
              f"max {err.max().item():.2e} | quant-noise median {qn.median().item():.2e} max {qn.max().item():.2e} | max|v| {s.item():.3f}")

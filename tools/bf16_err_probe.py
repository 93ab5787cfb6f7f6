"""Error distribution of the bf16 (tcgen05) value rollout against the quantised oracle, on the oracle's own CEM action rows
(iteration 0..n of a plan) and on uniform random rows.  python tools/bf16_err_probe.py --workload cfg1 --seed 0 [--lib variant.so]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg1"); ap.add_argument("--seed", type=int, default=0); ap.add_argument("--lib", default=None)
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
from tdmpc_b200 import _lib
if a.lib: _lib.LIB_PATH = os.path.abspath(a.lib)
from oracle import tdmpc_oracle as O
from tests import synth, gpu_util as gu
cfg = synth.make_cfg(a.workload); sd = synth.make_told_weights(cfg, a.seed)
pl = gu.make_planner(cfg, sd, a.precision)
obs = synth.make_obs(cfg, a.seed * 100); nz = synth.make_noise(cfg, a.seed * 100)
quant = a.precision == "bf16"
_, _, _, taps = gu.oracle_plan(sd, cfg, obs, nz, t0=True, prev_mean=None, quant=quant)
z0 = O.told_h(sd, torch.from_numpy(obs).unsqueeze(0))
def show(tag, v_g, v_o):
    err = (v_g - v_o).abs() / v_o.abs().max()
    print(f"{os.path.basename(a.lib or 'default')} {a.workload} seed {a.seed} {tag}: median {err.median().item():.2e} p75 {err.quantile(0.75).item():.2e} "
          f"p90 {err.quantile(0.9).item():.2e} p99 {err.quantile(0.99).item():.2e} max {err.max().item():.2e} | value range [{v_o.min().item():.3f}, {v_o.max().item():.3f}]")
for i in range(a.iters):
    acts = taps["actions"][i]
    v_g, _ = pl.estimate_value(z0.cuda(), acts.permute(1, 0, 2).unsqueeze(0).contiguous().cuda(), discount=cfg.discount, min_std=cfg.min_std,
                               noise_term=nz["term"][i].unsqueeze(0))
    show(f"cem iter {i}", v_g[0].cpu(), taps["values"][i])
rs = np.random.RandomState(a.seed + 7)
H, N, P, T = synth.plan_dims(cfg)
acts = torch.from_numpy(rs.uniform(-1, 1, (H, T, cfg.action_dim)).astype(np.float32))
eps = torch.from_numpy(rs.standard_normal((T, cfg.action_dim)).astype(np.float32))
v_o, _ = O.estimate_value(sd, z0.repeat(T, 1), acts, H, cfg.discount, cfg.min_std, eps, quant)
v_g, _ = pl.estimate_value(z0.cuda(), acts.permute(1, 0, 2).unsqueeze(0).contiguous(), discount=cfg.discount, min_std=cfg.min_std, noise_term=eps.unsqueeze(0))
show("uniform rows", v_g[0].cpu(), v_o.squeeze(1))

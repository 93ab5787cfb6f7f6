"""Deterministic synthetic inputs shared by the golden-fixture generator, the parity tests,
smoke() and bench.py.  Everything comes from numpy's legacy MT19937 RandomState so the same
seed gives the same bytes in the build container and on the GPU box (the fixtures under
tests/golden/ therefore store only OUTPUTS of the reference, not weights or noise).
"""
from types import SimpleNamespace

import numpy as np
import torch


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def make_cfg(name="cfg1", **over):
    """Planner/model scalars of the BASELINE.json configs (explicit, never read from yaml)."""
    base = dict(
        modality="state", device="cpu", enc_dim=256, mlp_dim=512, latent_dim=50, hidden_dim=128,
        horizon=5, num_samples=512, num_elites=64, iterations=6, mixture_coef=0.05,
        min_std=0.05, temperature=0.5, momentum=0.1, discount=0.99, seed_steps=5000,
        std_schedule="linear(0.5, 0.05, 25000, 0)", horizon_schedule="linear(2, 5, 25000, 0)",
        regularization_schedule="linear(0.05, 0.5, 1, 5000)", lr=1e-3,
        factor_decrease_num=1.25, shift_elites_over_time=True, fraction_elites_reused=0.25,
        keep_previous_elites=True, noise_beta=2.5, normalize=False, norm_type="ln", norm_cell=True,
        img_size=84, frame_stack=1, num_channels=32,
    )
    dims = {
        "cfg1": dict(obs_shape=(5,), action_dim=1),                       # cartpole-swingup
        "cfg2": dict(obs_shape=(223,), action_dim=38, latent_dim=100, iterations=8),  # dog-run (cfgs/tasks/dog.yaml)
        "cfg2s": dict(obs_shape=(223,), action_dim=38),                   # dog-run with L=50, I=6
        "cfg3": dict(obs_shape=(18,), action_dim=4),                      # quadrotor xyz_vxyz_R_omega
        "cfg4": dict(obs_shape=(18,), action_dim=4, horizon=10, num_samples=1024, num_elites=32,
                     mixture_coef=0.5, horizon_schedule="linear(2, 10, 25000, 0)", normalize=True),
        "tiny": dict(obs_shape=(7,), action_dim=3, latent_dim=24, horizon=3, num_samples=96,
                     num_elites=16, iterations=3, mixture_coef=0.25, horizon_schedule="3"),
    }[name]
    base.update(dims)
    base.update(over)
    return SimpleNamespace(**base)


def make_told_weights(cfg, seed=0, enc_ln=False):
    """Random TOLD parameters keyed like the reference state_dict (tdmpc.py:12-23).
    Non-degenerate on purpose: the reference zero-initialises the reward/Q heads
    (tdmpc.py:21-23), which makes every trajectory value 0 -- useless for parity."""
    rs = np.random.RandomState(seed)
    D, A, L, M, E = cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim
    sd = {}

    def lin(key, out, inp, gain=1.0, bias_std=0.05):
        sd[key + ".weight"] = _t(rs.standard_normal((out, inp)) * (gain / np.sqrt(inp)))
        sd[key + ".bias"] = _t(rs.standard_normal(out) * bias_std)

    def ln(key, n):
        sd[key + ".weight"] = _t(1.0 + 0.1 * rs.standard_normal(n))
        sd[key + ".bias"] = _t(0.1 * rs.standard_normal(n))

    if enc_ln:
        lin("_encoder.0", E, D); ln("_encoder.1", E); lin("_encoder.3", L, E)
    else:
        lin("_encoder.0", E, D); lin("_encoder.2", L, E)
    for name, out in (("_dynamics", L), ("_reward", 1), ("_pi", A)):
        inp = L if name == "_pi" else L + A
        lin(name + ".0", M, inp, gain=1.4); lin(name + ".2", M, M, gain=1.4); lin(name + ".4", out, M)
    for name in ("_Q1", "_Q2"):
        lin(name + ".0", M, L + A); ln(name + ".1", M)
        lin(name + ".3", M, M); ln(name + ".4", M)
        lin(name + ".6", 1, M, gain=2.0)
    return sd


def make_obs(cfg, seed=0, batch=None):
    rs = np.random.RandomState(1000 + seed)
    shape = (cfg.obs_shape[0],) if batch is None else (batch, cfg.obs_shape[0])
    return rs.standard_normal(shape).astype(np.float32)


def plan_dims(cfg, step=10 ** 6):
    from oracle.tdmpc_oracle import linear_schedule
    H = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    N = cfg.num_samples
    P = int(cfg.mixture_coef * N)
    return H, N, P, N + P


def make_noise(cfg, seed=0, step=10 ** 6, batch=None):
    """Standard-normal draws in the reference's consumption order."""
    rs = np.random.RandomState(2000 + seed)
    H, N, P, T = plan_dims(cfg, step)
    A, I = cfg.action_dim, cfg.iterations
    lead = () if batch is None else (batch,)
    n = {
        "pi": _t(rs.standard_normal(lead + (H, P, A))),
        "cem": _t(rs.standard_normal(lead + (I, H, N, A))),
        "term": _t(rs.standard_normal(lead + (I, T, A))),
        "out": _t(rs.standard_normal(lead + (A,))),
        "seed_uniform": _t(rs.uniform(-1, 1, lead + (A,))),
    }
    u = rs.random_sample(lead if lead else None)
    n["u"] = float(u) if batch is None else np.asarray(u, dtype=np.float64)
    return n

"""Shared helpers of the GPU parity tests (imported only under -m gpu)."""
import numpy as np
import torch

from oracle import tdmpc_oracle as O
from tests import synth


def make_planner(cfg, sd, precision):
    from tdmpc_b200 import Planner
    pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                 enc_layernorm="_encoder.3.weight" in sd, precision=precision)
    pl.sync_weights({k: v.cuda() for k, v in sd.items()})
    return pl


def plan_kwargs(cfg, step=10 ** 6, std_floor=None):
    H, N, P, T = synth.plan_dims(cfg, step)
    return dict(horizon=H, num_samples=N, num_pi_trajs=P, num_elites=cfg.num_elites, iterations=cfg.iterations,
                discount=cfg.discount, temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std,
                std_floor=cfg.min_std if std_floor is None else std_floor)


def batch_noise(noises):
    """list of per-problem reference-layout noise dicts -> ABI layout with leading B."""
    out = {k: torch.stack([n[k] for n in noises]).cuda() for k in ("pi", "cem", "term", "out")}
    out["u"] = np.array([n["u"] for n in noises], dtype=np.float64)
    return out


def oracle_plan(sd, cfg, obs, nz, *, t0, prev_mean, eval_mode=True, std_floor=None, quant=False, step=10 ** 6):
    return O.plan_tdmpc(sd, cfg, obs, eval_mode=eval_mode, step=step, t0=t0, prev_mean=prev_mean,
                        std_floor=cfg.min_std if std_floor is None else std_floor, noise=nz, quant=quant)


def compare_plan(res, b, oracle_out, cfg, tol, label=""):
    """Compare one problem of a GPU plan (with taps) against the oracle, iteration by iteration.
    Values / mean / std are checked while the elite sets agree; an elite-set difference is accepted
    only if every swapped trajectory sits within `tol` of the K-th value (unresolvable tie)."""
    a_o, m_o, mean_o, taps_o = oracle_out
    K = cfg.num_elites
    taps = res["taps"]
    diverged = False
    for i in range(cfg.iterations):
        v_o = taps_o["values"][i].numpy()
        T = v_o.shape[0]
        v_g = taps["values"][i, b, :T].cpu().numpy()
        scale = max(np.abs(v_o).max(), 1e-6)
        err = np.abs(v_g - v_o).max() / scale
        assert err <= tol, f"{label} iter {i}: value rel err {err:.3e} > {tol}"
        e_o = set(taps_o["elite_idx"][i].tolist())
        e_g = set(taps["elite_idx"][i, b].cpu().tolist())
        if e_o != e_g:
            kth = np.sort(v_o)[::-1][K - 1]
            for j in e_o ^ e_g:
                assert abs(v_o[j] - kth) <= 4 * tol * scale, f"{label} iter {i}: elite {j} differs and is not a near-tie"
            diverged = True
            break
        # order inside the elite list (sorted descending) where gaps are resolvable
        lo, lg = taps_o["elite_idx"][i].tolist(), taps["elite_idx"][i, b].cpu().tolist()
        for k in range(K):
            if lo[k] != lg[k]:
                assert abs(v_o[lo[k]] - v_o[lg[k]]) <= 4 * tol * scale, f"{label} iter {i}: elite order differs at rank {k}"
        np.testing.assert_allclose(taps["mean"][i, b].cpu().numpy(), taps_o["mean"][i].numpy(), rtol=0, atol=50 * tol,
                                   err_msg=f"{label} mean iter {i}")
        np.testing.assert_allclose(taps["std"][i, b].cpu().numpy(), taps_o["std"][i].numpy(), rtol=0, atol=50 * tol,
                                   err_msg=f"{label} std iter {i}")
    if not diverged:
        np.testing.assert_allclose(res["prev_mean"][b].cpu().numpy(), mean_o.numpy(), rtol=0, atol=50 * tol)
        np.testing.assert_allclose(res["action"][b].cpu().numpy(), a_o.numpy(), rtol=0, atol=50 * tol)
        mg = res["metrics"][b].cpu().numpy()
        assert abs(mg[0] - m_o["external_reward_mean"]) <= 50 * tol * max(1.0, abs(m_o["external_reward_mean"]))
        assert abs(mg[1] - m_o["current_std"]) <= 50 * tol
    return not diverged

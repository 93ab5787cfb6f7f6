"""Pins oracle/tdmpc_oracle.py against fixtures produced by the real reference planner
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import tdmpc_oracle as O
from tests import synth

CASES = [("tdmpc_tiny", "tiny", {}), ("tdmpc_tiny_train", "tiny", {"std_floor": 0.1}),
         ("tdmpc_cfg1", "cfg1", {}), ("tdmpc_cfg3", "cfg3", {}), ("tdmpc_cfg2s", "cfg2s", {})]

# fixtures were generated on another host CPU; MKL may pick different GEMM kernels there
TOL = 2e-5


@pytest.mark.parametrize("name,cfg_name,kw", CASES)
def test_oracle_matches_reference_fixture(golden_dir, name, cfg_name, kw):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, steps, eval_mode = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name)
    sd = synth.make_told_weights(cfg, seed)
    std_floor = kw.get("std_floor", cfg.min_std)
    step = 10 ** 6
    prev = None
    for s in range(steps):
        obs = synth.make_obs(cfg, seed * 100 + s)
        nz = synth.make_noise(cfg, seed * 100 + s, step)
        a, m, prev, taps = O.plan_tdmpc(sd, cfg, obs, eval_mode=bool(eval_mode), step=step, t0=(s == 0),
                                        prev_mean=prev, std_floor=std_floor, noise=nz)
        vals = torch.stack(taps["values"]).numpy()
        ref_vals = g[f"s{s}_values"]
        scale = np.abs(ref_vals).max()
        assert np.abs(vals - ref_vals).max() <= TOL * scale
        # elite sets: identical wherever the K-th / (K+1)-th gap is resolvable
        for i in range(cfg.iterations):
            srt = np.sort(ref_vals[i])[::-1]
            K = cfg.num_elites
            if srt[K - 1] - srt[K] > 4 * TOL * scale:
                assert set(taps["elite_idx"][i].tolist()) == set(g[f"s{s}_elite_idx"][i].tolist())
        np.testing.assert_allclose(prev.numpy(), g[f"s{s}_prev_mean"], rtol=0, atol=1e-4)
        np.testing.assert_allclose(a.numpy(), g[f"s{s}_action"], rtol=0, atol=1e-4)
        np.testing.assert_allclose([m["external_reward_mean"], m["current_std"]], g[f"s{s}_metrics"], rtol=1e-4, atol=1e-5)


def test_seed_steps_early_return():
    cfg = synth.make_cfg("tiny")
    sd = synth.make_told_weights(cfg, 0)
    nz = synth.make_noise(cfg, 0)
    a, m, prev, _ = O.plan_tdmpc(sd, cfg, synth.make_obs(cfg), eval_mode=False, step=10, t0=True,
                                 prev_mean=None, std_floor=0.05, noise=nz)
    assert torch.equal(a, nz["seed_uniform"]) and m == {"external_reward_mean": 0.0, "current_std": 0.0}


def test_linear_schedule():
    assert O.linear_schedule("0.25", 7) == 0.25
    assert O.linear_schedule("linear(0.5, 0.05, 25000, 0)", 0) == 0.5
    assert abs(O.linear_schedule("linear(0.5, 0.05, 25000, 0)", 12500) - 0.275) < 1e-12
    assert O.linear_schedule("linear(2, 5, 25000, 0)", 10 ** 6) == 5.0
    with pytest.raises(NotImplementedError):
        O.linear_schedule("cosine(1,2)", 0)


def test_categorical_pick_is_numpy_choice():
    for s in range(40):
        p = np.random.RandomState(s).dirichlet(np.ones(32)).astype(np.float32)
        np.random.seed(s); u = np.random.random_sample()
        np.random.seed(s); j = np.random.choice(np.arange(32), p=p)
        assert O.categorical_pick(torch.from_numpy(p).unsqueeze(1), u) == j


def test_flop_accounting_matches_survey():
    # cfg1 = 25.15 GFLOP / plan, 16,235 sample-steps
    f = O.flops_per_plan(D=5, A=1, L=50, M=512, E=256, H=5, rows_per_iter=[537] * 6, P=25)
    assert abs(f / 1e9 - 25.15) < 0.05
    assert O.sample_steps_per_plan(5, [537] * 6, 25) == 16235


@pytest.mark.parametrize("name,cfg_name", [("icem_tiny", "tiny_icem"), ("icem_cfg4", "cfg4")])
def test_icem_oracle_matches_reference_fixture(golden_dir, name, cfg_name):
    """oracle.plan_icem_mlp against what the real TdICemSimMlp.plan produced (colored noise, sample decay,
    shifted / carried elites, best-a): tests/golden/make_golden.py:run_icem_case."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, steps, _ = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name)
    sd = synth.make_told_weights(cfg, seed, enc_ln=bool(cfg.normalize))
    state = {"plan_horizon": 1, "std": cfg.min_std}
    for s in range(steps):
        obs = synth.make_obs(cfg, seed * 100 + s)
        have = state.get("_elite_actions") is not None
        sizes = O.icem_sizes(cfg, O.linear_schedule(cfg.regularization_schedule, 10 ** 6), have)
        assert np.array_equal(np.array(sizes), g[f"s{s}_sizes"])
        H = int(min(cfg.horizon, O.linear_schedule(cfg.horizon_schedule, 10 ** 6)))
        nz = synth.make_icem_noise(cfg, seed * 100 + s, sizes, H, have)
        a, m, taps = O.plan_icem_mlp(sd, cfg, obs, eval_mode=True, step=10 ** 6, t0=(s == 0), state=state, noise=nz)
        ref_vals = g[f"s{s}_values"]
        scale = np.abs(ref_vals).max()
        for i in range(cfg.iterations):
            v = taps["values"][i].numpy()
            assert np.abs(v - ref_vals[i, :v.shape[0]]).max() <= TOL * scale
        np.testing.assert_allclose(state["_prev_mean"].numpy(), g[f"s{s}_prev_mean"], atol=1e-4)
        np.testing.assert_allclose(state["_elite_actions"].numpy(), g[f"s{s}_elite_actions"], atol=1e-4)
        np.testing.assert_allclose(a.numpy(), g[f"s{s}_action"], atol=1e-4)


@pytest.mark.parametrize("name,cfg_name", [("drnn_tiny", "tiny_drnn"), ("drnn_cfg5", "cfg5")])
def test_drnn_oracle_matches_reference_fixture(golden_dir, name, cfg_name):
    """oracle.plan_icem_drnn against what the real TdICemSimDssm.plan produced (NormGRUCell + BatchNorm prior MLP
    dynamics, reward from the GRU state, hidden-state threading): tests/golden/make_golden.py:run_drnn_case."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, steps, _ = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name)
    sd = synth.make_drnn_weights(cfg, seed)
    state = {"plan_horizon": 1, "std": cfg.min_std}
    hidden = torch.zeros(1, cfg.hidden_dim)
    for s in range(steps):
        obs = synth.make_obs(cfg, seed * 100 + s)
        t0 = s == 0
        have = state.get("_elite_actions") is not None
        sizes = O.icem_drnn_sizes(cfg, O.linear_schedule(cfg.regularization_schedule, 10 ** 6), have, t0)
        assert np.array_equal(np.array(sizes), g[f"s{s}_sizes"])
        H = int(min(cfg.horizon, O.linear_schedule(cfg.horizon_schedule, 10 ** 6)))
        nz = synth.make_icem_noise(cfg, seed * 100 + s, sizes, H, have and not t0, single_beta=True)
        a, hidden, m, taps = O.plan_icem_drnn(sd, cfg, obs, hidden, eval_mode=True, step=10 ** 6, t0=t0, state=state, noise=nz)
        ref_vals = g[f"s{s}_values"]
        scale = np.abs(ref_vals).max()
        for i in range(cfg.iterations):
            v = taps["values"][i].numpy()
            assert np.abs(v - ref_vals[i, :v.shape[0]]).max() <= TOL * scale
        np.testing.assert_allclose(state["_prev_mean"].numpy(), g[f"s{s}_prev_mean"], atol=1e-4)
        np.testing.assert_allclose(state["_elite_actions"].numpy(), g[f"s{s}_elite_actions"], atol=1e-4)
        np.testing.assert_allclose(a.numpy(), g[f"s{s}_action"], atol=1e-4)
        np.testing.assert_allclose(hidden.numpy(), g[f"s{s}_hidden"], atol=1e-4)


def test_drnn_model_mirror_matches_oracle_step():
    """tdmpc_b200.DSSMRnn (the nn.Module mirror of the reference DSSM with GRU dynamics) has the reference's
    state_dict keys and its next() equals the oracle's dssm_next."""
    from tdmpc_b200 import DSSMRnn
    cfg = synth.make_cfg("tiny_drnn")
    sd = synth.make_drnn_weights(cfg, 1)
    m = DSSMRnn(cfg).eval()
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(k.startswith("_predictor") or k.endswith("num_batches_tracked") for k in missing.missing_keys)
    rs = np.random.RandomState(0)
    z = torch.from_numpy(rs.standard_normal((9, cfg.latent_dim)).astype(np.float32))
    a = torch.from_numpy(rs.uniform(-1, 1, (9, cfg.action_dim)).astype(np.float32))
    h = torch.from_numpy(rs.standard_normal((9, cfg.hidden_dim)).astype(np.float32))
    with torch.no_grad():
        z1, h1, r1 = m.next(z, a, h)
    z2, h2, r2 = O.dssm_next(sd, z, a, h)
    assert torch.allclose(z1, z2, atol=1e-6) and torch.allclose(h1, h2, atol=1e-6) and torch.allclose(r1, r2, atol=1e-6)


def test_colored_noise_matrix_is_the_generator():
    """tdmpc_b200.colored.synthesis_matrix (what the CUDA sampler multiplies white noise by) equals the oracle's
    restatement of colorednoise.powerlaw_psd_gaussian, both as a matrix and in distribution."""
    from oracle.colored_noise import powerlaw_psd_gaussian, synthesis_matrix as S_o
    from tdmpc_b200.colored import color_matrices, mix_split, synthesis_matrix
    for beta in (1.0, 2.5, 0.5):
        for n in (2, 5, 6, 10, 12):
            np.testing.assert_allclose(synthesis_matrix(beta, n), S_o(beta, n), atol=1e-12)
    x = powerlaw_psd_gaussian(2.5, (200000, 10), random_state=np.random.RandomState(0))
    S = synthesis_matrix(2.5, 10)
    assert np.abs(np.cov(x.T) - S @ S.T).max() < 2e-2
    m = color_matrices(10, 7, 2.5)
    assert m.shape == (3, 10, 10) and np.all(m[2, 7:] == 0) and np.all(m[2, :, 7:] == 0)
    assert mix_split(1024) == (341, 342, 341) and mix_split(819) == (273, 273, 273) and mix_split(335) == (111, 112, 112)


SIM_CASES = [("tdmpcsim_tiny", "tiny", dict(horizon=5, horizon_schedule="linear(2, 5, 25000, 0)", normalize=True)),
             ("tdmpcsim_cfg3", "cfg3", {})]


@pytest.mark.parametrize("name,cfg_name,over", SIM_CASES)
def test_oracle_matches_tdmpcsim_fixture(golden_dir, name, cfg_name, over):
    """plan_tdmpcsim against what the unmodified TDMPCSIM.plan (tdmpc_similarity.py:135-200) produced: plan_horizon moves
    only at t0, P follows regularization_schedule, bare action return, exploration noise (cfg3 case)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, ncalls, eval_mode = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name, batch_size=8, **over)
    sd = synth.make_told_weights(cfg, seed, enc_ln=bool(cfg.normalize))
    state = {"plan_horizon": 1, "std": 0.1}
    for s, (step, t0) in enumerate(g["plan"].tolist()):
        H, P = [int(x) for x in g[f"s{s}_hp"]]
        obs = synth.make_obs(cfg, seed * 100 + s)
        nz = synth.make_noise_hp(cfg, seed * 100 + s, H, P)
        a, taps = O.plan_tdmpcsim(sd, cfg, obs, eval_mode=bool(eval_mode), step=step, t0=bool(t0), state=state, noise=nz)
        assert state["plan_horizon"] == H and int(state["mixture_coef"] * cfg.num_samples) == P
        vals, ref_vals = torch.stack(taps["values"]).numpy(), g[f"s{s}_values"]
        scale = np.abs(ref_vals).max()
        assert np.abs(vals - ref_vals).max() <= TOL * scale
        for i in range(cfg.iterations):
            srt = np.sort(ref_vals[i])[::-1]
            if srt[cfg.num_elites - 1] - srt[cfg.num_elites] > 4 * TOL * scale:
                assert set(taps["elite_idx"][i].tolist()) == set(g[f"s{s}_elite_idx"][i].tolist())
        np.testing.assert_allclose(state["_prev_mean"].numpy(), g[f"s{s}_prev_mean"], rtol=0, atol=1e-4)
        np.testing.assert_allclose(a.numpy(), g[f"s{s}_action"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("name,cfg_name", [("update_tiny", "tiny"), ("update_cfg3", "cfg3"), ("update_cfg2s", "cfg2s")])
def test_oracle_update_forward_matches_reference_fixture(golden_dir, name, cfg_name):
    """oracle.update_forward against what the unmodified TDMPC.update (tdmpc.py:191-245) returned for the same replay batch:
    loss means, weighted loss, and the per-sample priorities it handed to ReplayBuffer.update_priorities."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, batch = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name, batch_size=batch)
    sd, sd_t = synth.make_told_weights(cfg, seed), synth.make_told_weights(cfg, seed + 50)
    o = O.update_forward(sd, sd_t, cfg, synth.make_update_batch(cfg, seed, batch))
    got = [float(o["consistency"].mean()), float(o["reward"].mean()), float(o["value"].mean()), float(o["total"].mean()), o["weighted_loss"]]
    np.testing.assert_allclose(got, g["metrics"][:5], rtol=2e-5)
    np.testing.assert_allclose(o["priority"].clamp(max=1e4).numpy(), g["priorities"], rtol=2e-5, atol=1e-6)
    assert tuple(o["zs"].shape) == (cfg.horizon + 1, batch, cfg.latent_dim)


def test_package_accounting_matches_oracle():
    from tdmpc_b200 import accounting as A
    for args in ((5, 1, 50, 512, 256, 5, [537] * 6, 25), (223, 38, 100, 512, 256, 5, [537] * 8, 25), (18, 4, 50, 512, 256, 10, [1544, 1236, 990], 512)):
        assert A.flops_per_plan(*args) == O.flops_per_plan(*args)
    assert A.sample_steps_per_plan(5, [537] * 6, 25) == O.sample_steps_per_plan(5, [537] * 6, 25) == 16235


@pytest.mark.parametrize("name,cfg_name,over", [("train_small", "tiny", dict(mlp_dim=64, enc_dim=32)), ("train_cfg3", "cfg3", {})])
def test_oracle_update_step_matches_reference_golden(golden_dir, name, cfg_name, over):
    """oracle.update_step (the whole TDMPC.update: backward, clip, both Adams, update_pi, ema) against what the unmodified
    reference method returned and left in its parameters over two consecutive calls (tests/golden/make_golden.py train)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, batch, steps = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name, batch_size=batch, **over)
    sd, sd_t, opt = synth.make_told_weights(cfg, seed), synth.make_told_weights(cfg, seed + 50), None
    for s in range(steps):
        b = synth.make_train_batch(cfg, seed + 7 * s, batch)
        m, sd, sd_t, _, opt = O.update_step(sd, sd_t, cfg, b, opt, update_target=(10 ** 6 + s) % cfg.update_freq == 0)
        got = [m[k] for k in ("consistency_loss", "reward_loss", "value_loss", "total_loss", "weighted_loss", "grad_norm", "pi_loss")]
        np.testing.assert_allclose(got, g["metrics"][s], rtol=1e-6, atol=1e-9)
    for tag, cur in (("model", sd), ("target", sd_t)):
        for k, v in cur.items():
            ref = g[f"{tag}/{k}"]
            if name == "train_small":
                np.testing.assert_allclose(v.numpy(), ref, rtol=0, atol=1e-7, err_msg=f"{tag}/{k}")
            else:
                f = v.double().flatten()
                idx = torch.linspace(0, f.numel() - 1, 16).long()
                np.testing.assert_allclose(np.concatenate([[f.sum().item(), (f * f).sum().item()], f[idx].numpy()]), ref, rtol=1e-9, atol=1e-9)

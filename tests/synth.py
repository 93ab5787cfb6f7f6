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
        img_size=84, frame_stack=1, num_channels=32, train_steps=500000, episode_length=1000, optim_id="adamw", pi_lr=1e-3,
        # training-step scalars (cfgs/default.yaml:32-44; rho 0.5 as upstream TD-MPC so that its powers are exercised)
        rho=0.5, consistency_coef=2.0, reward_coef=0.5, value_coef=0.1, grad_clip_norm=10, update_freq=2, tau=0.01,
    )
    dims = {
        "cfg1": dict(obs_shape=(5,), action_dim=1),                       # cartpole-swingup
        "cfg2": dict(obs_shape=(223,), action_dim=38, latent_dim=100, iterations=8),  # dog-run (cfgs/tasks/dog.yaml)
        "cfg2s": dict(obs_shape=(223,), action_dim=38),                   # dog-run with L=50, I=6
        "cfg3": dict(obs_shape=(18,), action_dim=4),                      # quadrotor xyz_vxyz_R_omega
        "k80": dict(obs_shape=(18,), action_dim=6, latent_dim=64, iterations=2),   # L+A = 70 -> 64 + 16: compact SECOND k-block
        # WIDE instantiations of the tcgen05 kernel (action_dim > 8): odd dims take the scalar fallback of the z' epilogue,
        # even dims the action-pair path -- with a swizzled second k-block (L+A = 90 -> 64 + 32) and a compact one (70 -> 64 + 16)
        "wide_odd": dict(obs_shape=(18,), action_dim=11, latent_dim=51, iterations=2),
        "wide_sw": dict(obs_shape=(18,), action_dim=30, latent_dim=60, iterations=2),
        "wide_ck": dict(obs_shape=(18,), action_dim=20, latent_dim=50, iterations=2),
        "cfg4": dict(obs_shape=(18,), action_dim=4, horizon=10, num_samples=1024, num_elites=32,
                     mixture_coef=0.5, horizon_schedule="linear(2, 10, 25000, 0)", normalize=True),
        "tiny": dict(obs_shape=(7,), action_dim=3, latent_dim=24, horizon=3, num_samples=96,
                     num_elites=16, iterations=3, mixture_coef=0.25, horizon_schedule="3"),
        "tiny_drnn": dict(obs_shape=(7,), action_dim=3, latent_dim=24, horizon=4, num_samples=120, hidden_dim=128,
                          num_elites=16, iterations=3, mixture_coef=0.5, horizon_schedule="4", normalize=True),
        "cfg5": dict(obs_shape=(18,), action_dim=4, horizon=10, num_samples=1024, num_elites=32, hidden_dim=128,
                     mixture_coef=0.5, horizon_schedule="linear(2, 10, 25000, 0)", normalize=True),
        "tiny_icem": dict(obs_shape=(7,), action_dim=3, latent_dim=24, horizon=4, num_samples=120,
                          num_elites=16, iterations=3, mixture_coef=0.5, horizon_schedule="4", normalize=True),
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


def make_drnn_weights(cfg, seed=0):
    """Random DSSM (GRU latent dynamics) parameters keyed like the reference state_dict
    (tdmpc_icem_similarity_drnn.py:14-36, models/gru_dyna.py:11-23, models/rnns.py:8-17).  LayerNorm / BatchNorm
    affine parameters and running statistics are perturbed so folding bugs show."""
    rs = np.random.RandomState(7000 + seed)
    D, A, L, M, E, Hd = cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, cfg.hidden_dim
    sd = {}

    def lin(key, out, inp, gain=1.0, bias=True, bias_std=0.05):
        sd[key + ".weight"] = _t(rs.standard_normal((out, inp)) * (gain / np.sqrt(inp)))
        if bias:
            sd[key + ".bias"] = _t(rs.standard_normal(out) * bias_std)

    def ln(key, n):
        sd[key + ".weight"] = _t(1.0 + 0.1 * rs.standard_normal(n))
        sd[key + ".bias"] = _t(0.1 * rs.standard_normal(n))

    lin("_encoder.0", E, D); ln("_encoder.1", E); lin("_encoder.3", L, E)
    lin("_dynamics.gru_cell.weight_ih", 3 * Hd, L + A, bias=False)
    lin("_dynamics.gru_cell.weight_hh", 3 * Hd, Hd, bias=False)
    for g in ("reset", "update", "newval"):
        ln(f"_dynamics.gru_cell.ln_{g}", Hd)
    lin("_dynamics.prior_mlp.0", M, Hd, gain=1.2)
    ln("_dynamics.prior_mlp.1", M)
    sd["_dynamics.prior_mlp.1.running_mean"] = _t(0.2 * rs.standard_normal(M))
    sd["_dynamics.prior_mlp.1.running_var"] = _t(np.exp(0.3 * rs.standard_normal(M)))
    lin("_dynamics.prior_mlp.3", L, M)
    lin("_reward.0", M, Hd, gain=1.4); lin("_reward.2", M, M, gain=1.4); lin("_reward.4", 1, M)
    lin("_pi.0", M, L, gain=1.4); lin("_pi.2", M, M, gain=1.4); lin("_pi.4", A, M)
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


def make_noise_hp(cfg, seed, H, P):
    """make_noise with the plan horizon and the number of policy trajectories given explicitly: TDMPCSIM takes them from
    `plan_horizon` (moved only at an episode start) and `regularization_schedule` (tdmpc_similarity.py:150-154)."""
    rs = np.random.RandomState(2500 + seed)
    A, I, N = cfg.action_dim, cfg.iterations, cfg.num_samples
    n = {
        "pi": _t(rs.standard_normal((H, P, A))),
        "cem": _t(rs.standard_normal((I, H, N, A))),
        "term": _t(rs.standard_normal((I, N + P, A))),
        "out": _t(rs.standard_normal((A,))),
        "seed_uniform": _t(rs.uniform(-1, 1, (A,))),
    }
    n["u"] = float(rs.random_sample())
    return n


def make_update_batch(cfg, seed, batch):
    """What ReplayBuffer.sample() hands to TDMPC.update (helper.py:506-534, latent_plan=True): obs [B,D], next_obses [H+1,B,D],
    action [H+1,B,A], reward [H+1,B,1], idxs [B], weights [B,1]; plus the standard-normal draws of pi() inside _td_target
    [H,B,A] (tdmpc.py:188)."""
    rs = np.random.RandomState(4000 + seed)
    H, D, A = cfg.horizon, cfg.obs_shape[0], cfg.action_dim
    return {"obs": _t(rs.standard_normal((batch, D))), "next_obses": _t(rs.standard_normal((H + 1, batch, D))),
            "action": _t(rs.uniform(-1, 1, (H + 1, batch, A))), "reward": _t(rs.standard_normal((H + 1, batch, 1))),
            "idxs": torch.from_numpy(rs.permutation(10 * batch)[:batch].astype(np.int64)),
            "weights": _t(rs.uniform(0.2, 1.0, (batch, 1))), "noise_pi": _t(rs.standard_normal((H, batch, A)))}


def make_train_batch(cfg, seed, batch):
    """make_update_batch plus the standard-normal draws of pi() inside update_pi (tdmpc.py:172): [H+1,B,A]."""
    b = make_update_batch(cfg, seed, batch)
    rs = np.random.RandomState(4500 + seed)
    b["noise_pi_update"] = _t(rs.standard_normal((cfg.horizon + 1, batch, cfg.action_dim)))
    return b


def make_icem_noise(cfg, seed, sizes, H, have_elites, single_beta=False):
    """Draws of one TdICemSimMlp.plan call in consumption order (tdmpc_icem_similarity_mlp.py:139-157,
    188-239): per iteration white randn + two colored blocks (series of length cfg.horizon, truncated to
    the plan horizon), the tail noise of the shifted elites, policy / terminal draws, pick, output noise.
    Returns the raw pieces (for injecting into the reference) and the assembled unit noise (oracle / ABI)."""
    from oracle.colored_noise import powerlaw_psd_gaussian
    from tdmpc_b200.colored import mix_split
    rs = np.random.RandomState(3000 + seed)
    A = cfg.action_dim
    P0 = sizes[0][2]
    n = {"pi": _t(rs.standard_normal((H, P0, A))), "white": [], "pink": [], "brown": [], "cem": [], "term": []}
    for (N, E, P) in sizes:
        if single_beta:      # DRNN agent: sample_action_sequence, one beta, series of length plan horizon
            c = powerlaw_psd_gaussian(cfg.noise_beta, (N, A, H), random_state=rs)
            n["pink"].append(c)
            n["cem"].append(torch.from_numpy(c).float().permute(2, 0, 1))
        else:
            nw, npk, nb = mix_split(N)
            w = _t(rs.standard_normal((H, nw, A)))
            pk = powerlaw_psd_gaussian(1.0, (npk, A, cfg.horizon), random_state=rs)
            br = powerlaw_psd_gaussian(2.5, (nb, A, cfg.horizon), random_state=rs)
            n["white"].append(w); n["pink"].append(pk); n["brown"].append(br)
            n["cem"].append(torch.cat([w, torch.from_numpy(pk).float().permute(2, 0, 1)[:H],
                                       torch.from_numpy(br).float().permute(2, 0, 1)[:H]], dim=1))
        n["term"].append(_t(rs.standard_normal((N + E + P, A))))
    E0 = sizes[0][1]
    if have_elites and E0 > 0:
        sh = powerlaw_psd_gaussian(cfg.noise_beta, (E0, A, H), random_state=rs) if cfg.noise_beta > 0 else rs.standard_normal((E0, A, H))
        n["shift_raw"] = sh
        n["shift"] = torch.from_numpy(sh).float().permute(2, 0, 1)
    n["out"] = _t(rs.standard_normal(A))
    n["seed_uniform"] = _t(rs.uniform(-1, 1, A))
    n["u"] = float(rs.random_sample())
    return n


def icem_noise_abi(n, sizes, H, A):
    """Pack make_icem_noise output into the C-ABI layouts (B = 1): cem [1,I,H,Nmax,A], term [1,I,Tmax,A] ..."""
    I = len(sizes)
    Nmax = max(s[0] for s in sizes); Tmax = max(sum(s) for s in sizes)
    cem = torch.zeros(1, I, H, Nmax, A); term = torch.zeros(1, I, Tmax, A)
    for i, (N, E, P) in enumerate(sizes):
        cem[0, i, :, :N] = n["cem"][i]
        term[0, i, :N + E + P] = n["term"][i]
    out = {"pi": n["pi"].unsqueeze(0), "cem": cem, "term": term, "out": n["out"].unsqueeze(0), "u": np.array([n["u"]])}
    if "shift" in n:
        out["shift"] = n["shift"].unsqueeze(0).contiguous()
    return out


# ---------------------------------------------------------------------------------------------
# Prioritised replay: deterministic buffer contents for the sampler tests.
def make_replay_cfg(name="small", **over):
    """cfg fields ReplayBuffer reads (helper.py:441-464,506-513)."""
    base = dict(modality="state", device="cpu", per_alpha=0.6, per_beta=0.4, horizon=5, env_horizon=5, frame_stack=1)
    base.update({
        # not yet full -> np.random.choice(replace=True)
        "small": dict(obs_shape=(6,), action_dim=2, episode_length=100, train_steps=4000, max_buffer_size=1000000,
                      batch_size=128, n_episodes=13),
        # wrapped around -> full -> replace=False, skewed priorities so that redraw rounds happen
        "full": dict(obs_shape=(6,), action_dim=2, episode_length=50, train_steps=1000000, max_buffer_size=1500,
                     batch_size=256, n_episodes=34),
    }[name])
    base.update(over)
    return SimpleNamespace(**base)


def make_replay_episodes(cfg, seed=0):
    """[(obs [ep+1, D], action [ep, A], reward [ep])] * cfg.n_episodes, float32."""
    rs = np.random.RandomState(1000 + seed)
    ep, D, A = cfg.episode_length, cfg.obs_shape[0], cfg.action_dim
    return [(rs.randn(ep + 1, D).astype(np.float32), rs.uniform(-1, 1, (ep, A)).astype(np.float32),
             rs.randn(ep).astype(np.float32)) for _ in range(cfg.n_episodes)]


def make_priority_updates(cfg, n_valid, seed=0, rounds=6, skew=3.0):
    """[(idxs int64 [B], priorities float32 [B, 1])]: what TDMPC.update feeds update_priorities (tdmpc.py:231);
    log-normal so a few transitions dominate.  Zero-priority (episode-tail) slots are left alone, as in training
    (they are never sampled, so never updated)."""
    rs = np.random.RandomState(2000 + seed)
    ep, H = cfg.episode_length, cfg.horizon
    out = []
    for _ in range(rounds):
        idxs = rs.randint(0, n_valid, size=cfg.batch_size)
        idxs = np.unique(idxs[(idxs % ep) < ep - H])      # distinct: a scatter with repeated indices has no defined winner on a GPU
        pr = np.exp(skew * rs.randn(len(idxs), 1)).astype(np.float32).clip(max=1e4)
        out.append((idxs.astype(np.int64), pr))
    return out


def make_uniforms(n, seed=0):
    """The float64 stream np.random.seed(seed); np.random.random_sample(n) would give."""
    return np.random.RandomState(seed).random_sample(n)

"""TEST INFRASTRUCTURE -- CPU oracle for the TD-MPC planning hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  The product path (tdmpc_b200/) never does and fails loudly when
its CUDA library is missing.

This is a straight-line restatement of the reference's planner in functional form
(weights = plain dict of tensors keyed like the reference state_dict).  It issues the same
aten CPU ops in the same order as the reference (F.linear -> addmm, F.elu, F.layer_norm,
tanh), so on one machine it reproduces the reference bit-for-bit; tests/golden/ holds
fixtures produced by the *real* reference code (tests/golden/make_golden.py, run in the
build container where /root/reference exists) and tests/test_oracle_golden.py pins this
file against them.  Parity status: **pinned by executing the reference itself** (the
reference ships no tests / golden vectors of its own -- SURVEY.md section 8c).

Reference lines followed (all under /root/reference/src):
  algorithm/tdmpc.py:30-50     TOLD.h / next / pi / Q
  algorithm/tdmpc.py:83-92     TDMPC.estimate_value
  algorithm/tdmpc.py:94-163    TDMPC.plan
  algorithm/tdmpc.py:180-228   TDMPC._td_target / update (forward half: loss terms, TD targets, priorities)
  algorithm/tdmpc.py:165-245   TDMPC.update_pi / update as a whole (update_step: autograd over the forward above, clip, Adam, ema)
  algorithm/helper.py:71-96    TruncatedNormal.sample
  algorithm/helper.py:119-133,156-176,197-201   enc / dmlab_enc_norm / mlp / q layer stacks
  algorithm/helper.py:639-652  linear_schedule
  algorithm/tdmpc_similarity.py:124-200   TDMPCSIM.estimate_value / plan
  algorithm/tdmpc_icem_similarity_mlp.py:115-265   iCEM estimate_value / sampling / plan
  algorithm/tdmpc_icem_similarity_drnn.py:56-59,124-146,168-272  DRNN next / plan
  models/gru_dyna.py:25-29, models/rnns.py:19-29, algorithm/helper.py:179-184  DGruDyna / NormGRUCell / mlp_norm

`quant=True` selects the "bf16-weight / fp32-accumulate" numerics of the tensor-core
kernel: every operand of a tensor-core GEMM (the Linear layers of dynamics, reward, pi,
Q1, Q2 except the scalar reward / Q heads, GRU gate GEMMs, prior MLP) is rounded to bf16
(round-to-nearest-even) -- weights AND layer inputs -- products are accumulated in fp32,
biases / LayerNorm / activations / the N=1 reward and Q heads / the encoder stay fp32.
"""
import math
import re

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- schedules


def linear_schedule(schdl, step):
    """helper.py:639-652."""
    try:
        return float(schdl)
    except ValueError:
        m = re.match(r"linear\((.+),(.+),(.+),(.+)\)", schdl)
        if m:
            init, final, duration, start = [float(g) for g in m.groups()]
            mix = np.clip((step - start) / duration, 0.0, 1.0)
            return (1.0 - mix) * init + mix * final
    raise NotImplementedError(schdl)


# ----------------------------------------------------------------------------- numerics


def _q(x, quant):
    """Round a GEMM operand to bf16 (RNE) and back to fp32 when quant is on."""
    return x.bfloat16().float() if quant else x


def _lin(sd, key, x, quant=False):
    """nn.Linear: y = x W^T + b with W [out, in]."""
    w = sd[key + ".weight"]
    b = sd.get(key + ".bias")
    return F.linear(_q(x, quant), _q(w, quant), b)


def _ln(sd, key, x, eps=1e-5):
    w = sd[key + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[key + ".bias"], eps)


def mlp3(sd, name, x, quant=False, head_quant=None):
    """helper.mlp (helper.py:169-176): Linear-ELU-Linear-ELU-Linear.

    head_quant: quantisation of the LAST layer; defaults to `quant` except for N=1 heads
    which the tensor-core kernel evaluates as an fp32 dot product in its epilogue.
    """
    h = F.elu(_lin(sd, name + ".0", x, quant))
    h = F.elu(_lin(sd, name + ".2", h, quant))
    if head_quant is None:
        head_quant = quant and sd[name + ".4.weight"].shape[0] > 1
    return _lin(sd, name + ".4", h, head_quant)


def q_head(sd, name, x, quant=False):
    """helper.q (helper.py:197-201): Linear-LN-Tanh-Linear-LN-ELU-Linear(->1)."""
    h = torch.tanh(_ln(sd, name + ".1", _lin(sd, name + ".0", x, quant)))
    h = F.elu(_ln(sd, name + ".4", _lin(sd, name + ".3", h, quant)))
    return _lin(sd, name + ".6", h, False)


# ----------------------------------------------------------------------------- TOLD


def told_h(sd, obs):
    """TOLD.h (tdmpc.py:30-32).  enc (helper.py:131-133) or, when the state_dict holds a
    LayerNorm at index 1, dmlab_enc_norm's state branch (helper.py:161-165).  Always fp32."""
    if "_encoder.3.weight" in sd:  # Linear, LayerNorm, ELU, Linear
        h = F.elu(_ln(sd, "_encoder.1", _lin(sd, "_encoder.0", obs)))
        return _lin(sd, "_encoder.3", h)
    return _lin(sd, "_encoder.2", F.elu(_lin(sd, "_encoder.0", obs)))


def told_next(sd, z, a, quant=False):
    """TOLD.next (tdmpc.py:34-37)."""
    x = torch.cat([z, a], dim=-1)
    return mlp3(sd, "_dynamics", x, quant), mlp3(sd, "_reward", x, quant)


def told_pi(sd, z, std=0.0, eps=None, quant=False, ste=False):
    """TOLD.pi (tdmpc.py:39-45) + TruncatedNormal.sample(clip=0.3) (helper.py:86-96).
    `eps` is the standard-normal draw the reference takes from _standard_normal."""
    # the policy head is a tensor-core GEMM in the kernel for every action_dim (N padded to 16), so its operands are
    # quantised even when action_dim == 1 (cartpole, BASELINE configs[0]); only the reward / Q heads are fp32 dot products
    mu = torch.tanh(mlp3(sd, "_pi", z, quant, head_quant=quant))
    if std > 0:
        e = eps * (torch.ones_like(mu) * std)
        e = torch.clamp(e, -0.3, 0.3)
        x = mu + e
        if ste:      # TruncatedNormal._clamp (helper.py:81-84): same value, gradient passed straight through (update_pi)
            return x - x.detach() + torch.clamp(x, -1.0 + 1e-6, 1.0 - 1e-6).detach()
        return torch.clamp(x, -1.0 + 1e-6, 1.0 - 1e-6)
    return mu


def told_Q(sd, z, a, quant=False):
    """TOLD.Q (tdmpc.py:47-50)."""
    x = torch.cat([z, a], dim=-1)
    return q_head(sd, "_Q1", x, quant), q_head(sd, "_Q2", x, quant)


# ----------------------------------------------------------------------------- DRNN (GRU latent dynamics)


def norm_gru_cell(sd, pre, x, h, quant=False):
    """NormGRUCell.forward (models/rnns.py:19-29); LN eps 1e-3; chunk order reset, update, newval."""
    gi = _lin(sd, pre + ".weight_ih", x, quant)
    gh = _lin(sd, pre + ".weight_hh", h, quant)
    ri, ui, ni = gi.chunk(3, 1)
    rh, uh, nh = gh.chunk(3, 1)
    reset = torch.sigmoid(_ln(sd, pre + ".ln_reset", ri + rh, 1e-3))
    update = torch.sigmoid(_ln(sd, pre + ".ln_update", ui + uh, 1e-3))
    newval = torch.tanh(_ln(sd, pre + ".ln_newval", ni + reset * nh, 1e-3))
    return update * newval + (1 - update) * h


def prior_mlp(sd, pre, h, quant=False):
    """helper.mlp_norm with BatchNorm1d in eval mode (helper.py:179-184)."""
    x = _lin(sd, pre + ".0", h, quant)
    x = F.batch_norm(x, sd[pre + ".1.running_mean"], sd[pre + ".1.running_var"],
                     sd[pre + ".1.weight"], sd[pre + ".1.bias"], False, 0.0, 1e-5)
    return _lin(sd, pre + ".3", F.elu(x), quant)


def dssm_next(sd, z, a, h_prev, quant=False):
    """DSSM.next (tdmpc_icem_similarity_drnn.py:56-59) + DGruDyna.forward (gru_dyna.py:25-29)."""
    x = torch.cat([z, a], dim=-1)
    h = norm_gru_cell(sd, "_dynamics.gru_cell", x, h_prev, quant)
    z_pred = prior_mlp(sd, "_dynamics.prior_mlp", h, quant)
    return z_pred, h, mlp3(sd, "_reward", h, quant)


# ----------------------------------------------------------------------------- estimate_value


def estimate_value(sd, z, actions, horizon, discount, min_std, eps_term, quant=False, hidden=None):
    """TDMPC.estimate_value (tdmpc.py:83-92).  actions [H,T,A]; eps_term [T,A].
    Returns (G [T,1] after nan_to_num, float mean of the last step's reward).
    With `hidden` (DRNN, tdmpc_icem_similarity_drnn.py:124-133) the GRU state is threaded."""
    G, disc = 0, 1
    reward = None
    for t in range(horizon):
        if hidden is None:
            z, reward = told_next(sd, z, actions[t], quant)
        else:
            z, hidden, reward = dssm_next(sd, z, actions[t], hidden, quant)
        G = G + disc * reward          # python-float discount -> rounded once to fp32
        disc *= discount
    q1, q2 = told_Q(sd, z, told_pi(sd, z, min_std, eps_term, quant), quant)
    G = G + disc * torch.min(q1, q2)
    return G.nan_to_num_(0), float(reward.mean().item())


# ----------------------------------------------------------------------------- selection / update


def topk_desc(value, k):
    """K largest of value [T], sorted descending.  torch.topk leaves tie order
    unspecified (tdmpc.py:138); the oracle (and the kernel) break ties by LOWER index."""
    v = value.detach().cpu().numpy()
    order = np.lexsort((np.arange(v.shape[0]), -v))
    return torch.from_numpy(order[:k].copy())


def cem_update(elite_value, elite_actions, mean, temperature, momentum, std_floor, std_max=2.0):
    """tdmpc.py:142-149.  elite_value [K,1], elite_actions [H,K,A], mean [H,A]."""
    max_value = elite_value.max(0)[0]
    score = torch.exp(temperature * (elite_value - max_value))
    score = score / score.sum(0)
    _mean = torch.sum(score.unsqueeze(0) * elite_actions, dim=1) / (score.sum(0) + 1e-9)
    _std = torch.sqrt(torch.sum(score.unsqueeze(0) * (elite_actions - _mean.unsqueeze(1)) ** 2, dim=1)
                      / (score.sum(0) + 1e-9))
    _std = _std.clamp_(std_floor, std_max)
    return momentum * mean + (1 - momentum) * _mean, _std, score


def categorical_pick(score, u):
    """np.random.choice(K, p=score) (tdmpc.py:152-153) with its single uniform draw `u`
    injected: numpy draws by inverse CDF on the float64 cumsum, side='right'."""
    p = score.squeeze(1).cpu().numpy().astype(np.float64)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


# ----------------------------------------------------------------------------- TDMPC.plan


def plan_tdmpc(sd, cfg, obs, *, eval_mode, step, t0, prev_mean, std_floor, noise, quant=False):
    """TDMPC.plan (tdmpc.py:94-163) with every random draw injected.

    cfg: object with horizon, horizon_schedule, mixture_coef, num_samples, num_elites,
         iterations, min_std, temperature, momentum, discount, action_dim, seed_steps.
    noise: dict  'pi'  [H,P,A]  -- _standard_normal draws of the policy rollout (Appendix B.1)
                 'cem' [I,H,N,A]-- torch.randn(H,N,A) per iteration
                 'term'[I,T,A]  -- terminal-policy draws inside estimate_value
                 'u'   float    -- uniform of np.random.choice
                 'out' [A]      -- exploration noise (not eval_mode)
                 'seed_uniform' [A] -- U(-1,1) action of the seed-step early return
    Returns (a [A], metrics, mean [H,A] (= new _prev_mean), taps).
    """
    metrics = {"external_reward_mean": 0.0, "current_std": 0.0}
    if step < cfg.seed_steps and not eval_mode:
        return noise["seed_uniform"].clone(), metrics, prev_mean, {}
    A = cfg.action_dim
    obs = torch.as_tensor(obs, dtype=torch.float32).unsqueeze(0)
    H = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    N, K = cfg.num_samples, cfg.num_elites
    P = int(cfg.mixture_coef * N)
    if P > 0:
        pi_actions = torch.empty(H, P, A)
        z = told_h(sd, obs).repeat(P, 1)
        for t in range(H):
            pi_actions[t] = told_pi(sd, z, cfg.min_std, noise["pi"][t], quant)
            z, _ = told_next(sd, z, pi_actions[t], quant)
    z = told_h(sd, obs).repeat(N + P, 1)
    mean = torch.zeros(H, A)
    std = 2 * torch.ones(H, A)
    if not t0 and prev_mean is not None:
        mean[:-1] = prev_mean[1:]
    taps = {"values": [], "elite_idx": [], "mean": [], "std": [], "z0": z[0].clone(), "actions": [], "mean_in": []}
    for i in range(cfg.iterations):
        actions = torch.clamp(mean.unsqueeze(1) + std.unsqueeze(1) * noise["cem"][i], -1, 1)
        if P > 0:
            actions = torch.cat([actions, pi_actions], dim=1)
        taps["actions"].append(actions.clone()); taps["mean_in"].append(mean.clone())      # teacher-forcing inputs (tests)
        value, reward_mean = estimate_value(sd, z, actions, H, cfg.discount, cfg.min_std,
                                            noise["term"][i], quant)
        elite_idxs = topk_desc(value.squeeze(1), K)
        elite_value, elite_actions = value[elite_idxs], actions[:, elite_idxs]
        mean, std, score = cem_update(elite_value, elite_actions, mean, cfg.temperature,
                                      cfg.momentum, std_floor)
        taps["values"].append(value.squeeze(1).clone()); taps["elite_idx"].append(elite_idxs.clone())
        taps["mean"].append(mean.clone()); taps["std"].append(std.clone())
    j = categorical_pick(score, noise["u"])
    a = elite_actions[0, j].clone()
    if not eval_mode:
        a = a + std[0] * noise["out"]
    metrics = {"current_std": std[0].mean().item(), "external_reward_mean": reward_mean}
    taps["pick"] = j
    taps["score"] = score.squeeze(1).clone()
    return a, metrics, mean, taps


# ----------------------------------------------------------------------------- TDMPCSIM.plan


def plan_tdmpcsim(sd, cfg, obs, *, eval_mode, step, t0, state, noise, quant=False, fine_tuning=False):
    """TDMPCSIM.plan (tdmpc_similarity.py:135-200) with every random draw injected -- the agent the quadrotor scripts
    build (train_dyna_episode.py:86,144).  Against TDMPC.plan: returns the bare action; `fine_tuning` disables the
    seed-step early return (:144); the horizon lives in state['plan_horizon'] and only moves when t0 (:151-152); the
    number of policy trajectories follows regularization_schedule (:153-154); estimate_value returns G only (:124-133).
    state: plan_horizon, std, _prev_mean.  noise: as plan_tdmpc with shapes for (plan_horizon, P).  Returns (a, taps)."""
    if step < cfg.seed_steps and not eval_mode and not fine_tuning:
        return noise["seed_uniform"].clone(), {}
    A = cfg.action_dim
    obs = torch.as_tensor(obs, dtype=torch.float32).unsqueeze(0)
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    if horizon != state["plan_horizon"] and t0:
        state["plan_horizon"] = horizon
    H = state["plan_horizon"]
    mix = linear_schedule(cfg.regularization_schedule, step)
    state["mixture_coef"] = mix
    N, K = cfg.num_samples, cfg.num_elites
    P = int(mix * N)
    if P > 0:
        pi_actions = torch.empty(H, P, A)
        z = told_h(sd, obs).repeat(P, 1)
        for t in range(H):
            pi_actions[t] = told_pi(sd, z, cfg.min_std, noise["pi"][t], quant)
            z, _ = told_next(sd, z, pi_actions[t], quant)
    z = told_h(sd, obs).repeat(N + P, 1)
    mean = torch.zeros(H, A)
    std = 2 * torch.ones(H, A)
    pm = state.get("_prev_mean")
    if not t0 and pm is not None:
        mean[:-1] = pm[1:]
    taps = {"values": [], "elite_idx": [], "mean": [], "std": []}
    for i in range(cfg.iterations):
        actions = torch.clamp(mean.unsqueeze(1) + std.unsqueeze(1) * noise["cem"][i], -1, 1)
        if P > 0:
            actions = torch.cat([actions, pi_actions], dim=1)
        value, _ = estimate_value(sd, z, actions, H, cfg.discount, cfg.min_std, noise["term"][i], quant)
        elite_idxs = topk_desc(value.squeeze(1), K)
        elite_value, elite_actions = value[elite_idxs], actions[:, elite_idxs]
        mean, std, score = cem_update(elite_value, elite_actions, mean, cfg.temperature, cfg.momentum, state["std"])
        taps["values"].append(value.squeeze(1).clone()); taps["elite_idx"].append(elite_idxs.clone())
        taps["mean"].append(mean.clone()); taps["std"].append(std.clone())
    j = categorical_pick(score, noise["u"])
    state["_prev_mean"] = mean
    a = elite_actions[0, j].clone()
    if not eval_mode:
        a = a + std[0] * noise["out"]
    taps["pick"] = j
    return a, taps


# ----------------------------------------------------------------------------- iCEM (MLP dynamics)


def icem_sizes(cfg, mixture_coef, have_elites):
    """Per-iteration (N_i, E_i, P_i) of tdmpc_icem_similarity_mlp.py:182-210."""
    out = []
    n = cfg.num_samples
    for i in range(cfg.iterations):
        if i > 0:
            n = max(2 * cfg.num_elites, int(n / cfg.factor_decrease_num))
        p = int(mixture_coef * n)
        have = have_elites or i > 0
        e = int(cfg.fraction_elites_reused * cfg.num_elites) if (cfg.fraction_elites_reused > 0 and have) else 0
        out.append((n, e, p))
    return out


def plan_icem_mlp(sd, cfg, obs, *, eval_mode, step, t0, state, noise, quant=False):
    """TdICemSimMlp.plan (tdmpc_icem_similarity_mlp.py:159-265), draws injected.

    state: dict carrying agent attributes across calls: plan_horizon, _prev_mean, _elite_actions, std.
    noise: 'pi' [H,P0,A]; 'cem' list over iterations of unit-variance noise [H,N_i,A]
           (already colored -- the 1/3 white + 1/3 pink + 1/3 brown mix of :139-157);
           'shift' [h,E,A] noise for the appended tail of shifted elites (:220-223);
           'term' list of [T_i,A]; 'u'; 'out' [A]; 'seed_uniform' [A].
    """
    metrics = {"external_reward_mean": 0.0, "current_std": 0.0}
    if step < cfg.seed_steps and not eval_mode:
        return noise["seed_uniform"].clone(), metrics, {}
    A = cfg.action_dim
    extend_horizon = False
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    if horizon != state["plan_horizon"] and t0:
        state["plan_horizon"] = horizon
        extend_horizon = True
    H = state["plan_horizon"]
    mix = linear_schedule(cfg.regularization_schedule, step)
    state["mixture_coef"] = mix
    mean = torch.zeros(H, A)
    std = 0.5 * torch.ones(H, A)
    pm = state.get("_prev_mean")
    if not t0 and pm is not None:
        mean[:-1] = pm[1:]
        mean[-1] = pm[-1]
    N = cfg.num_samples
    P = int(mix * N)
    assert P > 0
    obs = torch.as_tensor(obs, dtype=torch.float32).unsqueeze(0)
    z = told_h(sd, obs)
    pi_actions = torch.empty(H, P, A)
    zs = z.repeat(P, 1)
    for t in range(H):
        pi_actions[t] = told_pi(sd, zs, cfg.min_std, noise["pi"][t], quant)
        zs, _ = told_next(sd, zs, pi_actions[t], quant)
    taps = {"values": [], "elite_idx": [], "mean": [], "std": [], "sizes": []}
    for i in range(cfg.iterations):
        if i > 0:
            N = max(2 * cfg.num_elites, int(N / cfg.factor_decrease_num))
            P = int(mix * N)
        have = state.get("_elite_actions") is not None
        E = int(cfg.fraction_elites_reused * cfg.num_elites) if (cfg.fraction_elites_reused > 0 and have) else 0
        T = N + P + E
        zs_plan = z.repeat(T, 1)
        actions_sampled = torch.clamp(mean.unsqueeze(1) + std.unsqueeze(1) * noise["cem"][i], -1, 1)
        if i == cfg.iterations - 1:
            actions_sampled[:, 0] = mean
        if i == 0 and cfg.shift_elites_over_time and have:
            E = int(cfg.fraction_elites_reused * cfg.num_elites)
            reused = state["_elite_actions"][1:, :E]
            tail = torch.clamp(mean.unsqueeze(1) + std.unsqueeze(1) * noise["shift"], -1, 1)
            tail = tail[-2:] if extend_horizon else tail[-1:]
            reused = torch.cat([reused, tail], dim=0)
        if i > 0 and cfg.keep_previous_elites:
            reused = state["_elite_actions"][:, :E]
        if E > 0:
            actions = torch.cat([actions_sampled, reused, pi_actions[:, :P]], dim=1)
        else:
            actions = torch.cat([actions_sampled, pi_actions[:, :P]], dim=1)
        taps.setdefault("actions", []).append(actions.clone()); taps.setdefault("mean_in", []).append(mean.clone())
        value, reward_mean = estimate_value(sd, zs_plan, actions, H, cfg.discount, cfg.min_std,
                                            noise["term"][i], quant)
        elite_idxs = topk_desc(value.squeeze(1), cfg.num_elites)
        elite_value, elite_actions = value[elite_idxs], actions[:, elite_idxs]
        state["_elite_actions"] = elite_actions
        mean, std, score = cem_update(elite_value, elite_actions, mean, cfg.temperature,
                                      cfg.momentum, state["std"])
        taps["values"].append(value.squeeze(1).clone()); taps["elite_idx"].append(elite_idxs.clone())
        taps["mean"].append(mean.clone()); taps["std"].append(std.clone()); taps["sizes"].append((N, E, P))
    j = categorical_pick(score, noise["u"])
    state["_prev_mean"] = mean
    a = elite_actions[0, j].clone()
    if not eval_mode:
        a = a + std[0] * noise["out"]
    metrics = {"external_reward_mean": reward_mean, "current_std": std[0].mean().item()}
    taps["pick"] = j
    taps["score"] = score.squeeze(1).clone()
    return a, metrics, taps


# ----------------------------------------------------------------------------- iCEM (DRNN / GRU dynamics)


def icem_drnn_sizes(cfg, mixture_coef, have_elites, t0):
    """Per-iteration (N_i, E_i, P_i) of TdICemSimDssm.plan (tdmpc_icem_similarity_drnn.py:203-214): as the MLP
    variant, but elite reuse is gated on `not t0` in EVERY iteration."""
    out = []
    n = cfg.num_samples
    for i in range(cfg.iterations):
        if i > 0:
            n = max(2 * cfg.num_elites, int(n / cfg.factor_decrease_num))
        p = int(mixture_coef * n)
        have = (have_elites or i > 0) and not t0
        e = int(cfg.fraction_elites_reused * cfg.num_elites) if (cfg.fraction_elites_reused > 0 and have) else 0
        out.append((n, e, p))
    return out


def plan_icem_drnn(sd, cfg, obs, hidden, *, eval_mode, step, t0, state, noise, quant=False):
    """TdICemSimDssm.plan (tdmpc_icem_similarity_drnn.py:168-272), draws injected.
    hidden [1, Hd] in; returns (a, hidden' [1, Hd], metrics, taps).  noise as plan_icem_mlp ('cem' already
    colored with the single cfg.noise_beta of sample_action_sequence, :135-146)."""
    metrics = {"external_reward_mean": 0.0, "current_std": 0.0}
    if step < cfg.seed_steps and not eval_mode:
        return noise["seed_uniform"].clone(), None, metrics, {}
    A = cfg.action_dim
    extend_horizon = False
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    if horizon != state["plan_horizon"] and t0:
        state["plan_horizon"] = horizon
        extend_horizon = True
    H = state["plan_horizon"]
    mix = linear_schedule(cfg.regularization_schedule, step)
    state["mixture_coef"] = mix
    mean = torch.zeros(H, A)
    std = 0.5 * torch.ones(H, A)
    pm = state.get("_prev_mean")
    if not t0 and pm is not None:
        mean[:-1] = pm[1:]
        mean[-1] = pm[-1]
    N = cfg.num_samples
    P = int(mix * N)
    assert P > 0
    obs = torch.as_tensor(obs, dtype=torch.float32).unsqueeze(0)
    z = told_h(sd, obs)
    pi_actions = torch.empty(H, P, A)
    zs, hs = z.repeat(P, 1), hidden.repeat(P, 1)
    for t in range(H):
        pi_actions[t] = told_pi(sd, zs, cfg.min_std, noise["pi"][t], quant)
        zs, hs, _ = dssm_next(sd, zs, pi_actions[t], hs, quant)
    taps = {"values": [], "elite_idx": [], "mean": [], "std": [], "sizes": []}
    for i in range(cfg.iterations):
        if i > 0:
            N = max(2 * cfg.num_elites, int(N / cfg.factor_decrease_num))
            P = int(mix * N)
        have = state.get("_elite_actions") is not None
        E = int(cfg.fraction_elites_reused * cfg.num_elites) if (cfg.fraction_elites_reused > 0 and have and not t0) else 0
        T = N + P + E
        actions_sampled = torch.clamp(mean.unsqueeze(1) + std.unsqueeze(1) * noise["cem"][i], -1, 1)
        if i == cfg.iterations - 1:
            actions_sampled[:, 0] = mean
        if i == 0 and cfg.shift_elites_over_time and have and not t0:
            reused = state["_elite_actions"][1:, :E]
            tail = torch.clamp(mean.unsqueeze(1) + std.unsqueeze(1) * noise["shift"], -1, 1)
            tail = tail[-2:] if extend_horizon else tail[-1:]
            reused = torch.cat([reused, tail], dim=0)
        if i > 0 and cfg.keep_previous_elites:
            reused = state["_elite_actions"][:, :E]
        actions = torch.cat([actions_sampled, reused, pi_actions[:, :P]], dim=1) if E > 0 else torch.cat([actions_sampled, pi_actions[:, :P]], dim=1)
        taps.setdefault("actions", []).append(actions.clone()); taps.setdefault("mean_in", []).append(mean.clone())
        value, reward_mean = estimate_value(sd, z.repeat(T, 1), actions, H, cfg.discount, cfg.min_std, noise["term"][i], quant,
                                            hidden=hidden.repeat(T, 1))
        elite_idxs = topk_desc(value.squeeze(1), cfg.num_elites)
        elite_value, elite_actions = value[elite_idxs], actions[:, elite_idxs]
        state["_elite_actions"] = elite_actions
        mean, std, score = cem_update(elite_value, elite_actions, mean, cfg.temperature, cfg.momentum, state["std"])
        taps["values"].append(value.squeeze(1).clone()); taps["elite_idx"].append(elite_idxs.clone())
        taps["mean"].append(mean.clone()); taps["std"].append(std.clone()); taps["sizes"].append((N, E, P))
    j = categorical_pick(score, noise["u"])
    state["_prev_mean"] = mean
    a = elite_actions[0, j].clone()
    if not eval_mode:
        a = a + std[0] * noise["out"]
    _, hidden_next, _ = dssm_next(sd, z[0:1], a.unsqueeze(0), hidden, quant)
    metrics = {"external_reward_mean": reward_mean, "current_std": std[0].mean().item()}
    taps["pick"] = j
    taps["score"] = score.squeeze(1).clone()
    return a, hidden_next, metrics, taps


# ----------------------------------------------------------------------------- TDMPC.update, forward half


def update_forward(sd, sd_target, cfg, batch):
    """Forward of TDMPC.update (tdmpc.py:191-228) with _td_target (tdmpc.py:180-189), fp32, for a replay batch as
    tests/synth.make_update_batch builds it (state observations: h.RandomShiftsAug is the identity, helper.py:256-262).
    Returns the per-sample loss tensors [B,1], total / weighted loss and the latents zs the policy update consumes."""
    obs, next_obses, action, reward = batch["obs"], batch["next_obses"], batch["action"], batch["reward"]
    z = told_h(sd, obs)
    zs = [z]
    consistency_loss, reward_loss, value_loss, priority_loss = 0, 0, 0, 0
    taps = {"td_target": [], "q1": [], "q2": [], "reward_pred": []}
    for t in range(cfg.horizon):
        Q1, Q2 = told_Q(sd, z, action[t])
        z, reward_pred = told_next(sd, z, action[t])
        next_obs = next_obses[t]
        with torch.no_grad():                                         # tdmpc.py:207-210 (matters when update_step differentiates this)
            next_z = told_h(sd_target, next_obs)
            nz_on = told_h(sd, next_obs)                              # _td_target: ONLINE encoder and policy, TARGET Q heads
            td_target = reward[t] + cfg.discount * torch.min(*told_Q(sd_target, nz_on, told_pi(sd, nz_on, cfg.min_std, batch["noise_pi"][t])))
        zs.append(z)
        rho = cfg.rho ** t
        consistency_loss = consistency_loss + rho * torch.mean(F.mse_loss(z, next_z, reduction="none"), dim=1, keepdim=True)
        reward_loss = reward_loss + rho * F.mse_loss(reward_pred, reward[t], reduction="none")
        value_loss = value_loss + rho * (F.mse_loss(Q1, td_target, reduction="none") + F.mse_loss(Q2, td_target, reduction="none"))
        priority_loss = priority_loss + rho * (F.l1_loss(Q1, td_target, reduction="none") + F.l1_loss(Q2, td_target, reduction="none"))
        for k, v in (("td_target", td_target), ("q1", Q1), ("q2", Q2), ("reward_pred", reward_pred)):
            taps[k].append(v.squeeze(1).clone())
    total_loss = cfg.consistency_coef * consistency_loss.clamp(max=1e4) + cfg.reward_coef * reward_loss.clamp(max=1e4) + \
        cfg.value_coef * value_loss.clamp(max=1e4)
    weighted_loss = (total_loss * batch["weights"]).mean()
    return {"consistency": consistency_loss, "reward": reward_loss, "value": value_loss, "priority": priority_loss,
            "total": total_loss, "weighted_loss": float(weighted_loss.detach()), "zs": torch.stack(zs),
            "taps": {k: torch.stack(v) for k, v in taps.items()}}


# ----------------------------------------------------------------------------- TDMPC.update, the whole step

PI_KEYS = tuple(f"_pi.{i}.{s}" for i in (0, 2, 4) for s in ("weight", "bias"))


def _clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (norm_type 2, error_if_nonfinite False): returns the total norm, scales in place."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def _adam_step(params, grads, state, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam, single-tensor path (torch/optim/adam.py::_single_tensor_adam; no amsgrad / weight decay / maximize).
    `state[k]` = (exp_avg, exp_avg_sq), created as zeros on first use; `step` is the count AFTER this call."""
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    for k, p in params.items():
        g = grads[k]
        if k not in state:
            state[k] = (torch.zeros_like(p), torch.zeros_like(p))
        m, v = state[k]
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def update_step(sd, sd_target, cfg, batch, opt=None, update_target=True):
    """TDMPC.update (tdmpc.py:191-245) as a whole: forward (update_forward above), backward of weighted_loss / horizon,
    clip_grad_norm_ + Adam over the model (the policy block has no gradient there), update_pi (tdmpc.py:165-178) on the
    detached latents with the Q heads that were just stepped, its own clip + Adam, and helper.ema (helper.py:48-52).
    `batch` additionally holds "noise_pi_update" [H+1,B,A], the draws of pi() inside update_pi.  `opt` carries the optimiser
    state between calls: {"state": {}, "pi_state": {}, "step": 0, "pi_step": 0}.  Nothing is modified in place: returns
    (metrics, new_sd, new_sd_target, grads, opt) with `grads` the CLIPPED gradients (the policy block: update_pi's)."""
    opt = {"state": {}, "pi_state": {}, "step": 0, "pi_step": 0} if opt is None else \
        {"state": {k: (a.clone(), b.clone()) for k, (a, b) in opt["state"].items()},
         "pi_state": {k: (a.clone(), b.clone()) for k, (a, b) in opt["pi_state"].items()}, "step": opt["step"], "pi_step": opt["pi_step"]}
    P = {k: v.clone().requires_grad_(k not in PI_KEYS) for k, v in sd.items()}
    T = {k: v.clone() for k, v in sd_target.items()}
    with torch.enable_grad():
        o = update_forward(P, T, cfg, batch)
        weighted = (o["total"] * batch["weights"]).mean()
        (weighted * (1 / cfg.horizon)).backward()              # register_hook(grad * (1 / horizon)), tdmpc.py:224
    model_keys = [k for k in sd if k not in PI_KEYS]
    grads = {k: P[k].grad for k in model_keys}
    grad_norm = _clip_grad_norm([grads[k] for k in model_keys], cfg.grad_clip_norm)
    new = {k: v.detach().clone() for k, v in P.items()}
    opt["step"] += 1
    _adam_step({k: new[k] for k in model_keys}, grads, opt["state"], opt["step"], cfg.lr)
    # update_pi: Q heads frozen (track_q_grad(False)), latents detached
    Q = {k: v.clone() for k, v in new.items()}
    for k in PI_KEYS:
        Q[k] = new[k].clone().requires_grad_(True)
    zs = o["zs"].detach()
    with torch.enable_grad():
        pi_loss = 0
        for t in range(zs.shape[0]):
            a = told_pi(Q, zs[t], cfg.min_std, batch["noise_pi_update"][t], ste=True)
            q = torch.min(*told_Q(Q, zs[t], a))
            pi_loss = pi_loss + (-q.mean() * (cfg.rho ** t))
        pi_loss.backward()
    pi_grads = {k: Q[k].grad for k in PI_KEYS}
    pi_norm = _clip_grad_norm([pi_grads[k] for k in PI_KEYS], cfg.grad_clip_norm)
    opt["pi_step"] += 1
    _adam_step({k: new[k] for k in PI_KEYS}, pi_grads, opt["pi_state"], opt["pi_step"], cfg.lr)
    if update_target:
        for k in T:
            T[k].lerp_(new[k], cfg.tau)
    grads.update(pi_grads)
    f = lambda x: float(x.detach())
    metrics = {"consistency_loss": f(o["consistency"].mean()), "reward_loss": f(o["reward"].mean()),
               "value_loss": f(o["value"].mean()), "pi_loss": f(pi_loss), "total_loss": f(o["total"].mean()),
               "weighted_loss": f(weighted), "grad_norm": f(grad_norm), "pi_grad_norm": f(pi_norm),
               "priorities": o["priority"].detach().clamp(max=1e4)}
    return metrics, new, T, {k: g.detach() for k, g in grads.items()}, opt


# ----------------------------------------------------------------------------- flop / step accounting


def flops_per_plan(D, A, L, M, E, H, rows_per_iter, P):
    """Algorithmic FLOPs (2*MAC) of one plan: SURVEY.md section 8(d)."""
    f_dyn = 2 * ((L + A) * M + M * M + M * L)
    f_rew = 2 * ((L + A) * M + M * M + M)
    f_pi = 2 * (L * M + M * M + M * A)
    f_q = f_rew
    f_enc = 2 * (D * E + E * L)
    tot = f_enc + P * H * (f_pi + f_dyn)
    for T in rows_per_iter:
        tot += T * (H * (f_dyn + f_rew) + f_pi + 2 * f_q)
    return tot


def sample_steps_per_plan(H, rows_per_iter, P):
    return sum(T * H for T in rows_per_iter) + P * H

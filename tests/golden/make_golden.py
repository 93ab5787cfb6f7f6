"""Generate tests/golden/*.npz by running the UNMODIFIED reference planner
(/root/reference/src/algorithm/tdmpc.py: TDMPC.plan) on CPU with injected random draws.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

What is stored: only the reference's OUTPUTS (per-iteration trajectory values, top-k
indices, final mean, action, metrics).  Inputs (weights, observations, noise) are
regenerated from seeds by tests/synth.py, which is importable everywhere.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402

ref_shims.install()

from tests import synth  # noqa: E402
from oracle import tdmpc_oracle as O  # noqa: E402

import algorithm.helper as ref_h  # noqa: E402
from algorithm.tdmpc import TDMPC  # noqa: E402


class Injector:
    """Replaces torch.randn / helper._standard_normal / np.random.choice by queues of
    pre-drawn values, in the order the reference consumes them."""

    def __init__(self):
        self.randn_q, self.stdn_q, self.u = [], [], None
        self.topk_log, self.value_log = [], []

    def __enter__(self):
        self._randn, self._stdn, self._choice, self._topk = torch.randn, ref_h._standard_normal, np.random.choice, torch.topk

        def randn(*shape, **kw):
            x = self.randn_q.pop(0)
            assert tuple(x.shape) == tuple(shape), (x.shape, shape)
            return x.clone()

        def stdn(shape, dtype=None, device=None):
            x = self.stdn_q.pop(0)
            assert tuple(x.shape) == tuple(shape), (x.shape, shape)
            return x.clone()

        def choice(arr, p=None):
            # numpy's algorithm for choice(p=...): inverse CDF on the float64 cumsum.
            cdf = np.cumsum(np.asarray(p, dtype=np.float64)); cdf /= cdf[-1]
            return arr[int(np.searchsorted(cdf, self.u, side="right"))]

        def topk(x, k, dim=0):
            r = self._topk(x, k, dim=dim)
            self.topk_log.append(r.indices.clone())
            return r

        torch.randn, ref_h._standard_normal, np.random.choice, torch.topk = randn, stdn, choice, topk
        return self

    def __exit__(self, *a):
        torch.randn, ref_h._standard_normal, np.random.choice, torch.topk = self._randn, self._stdn, self._choice, self._topk


def check_choice_replica():
    """The injected categorical pick must be numpy's own algorithm."""
    for s in range(50):
        p = np.random.RandomState(s).dirichlet(np.ones(64)).astype(np.float32)
        np.random.seed(s); u = np.random.random_sample()
        np.random.seed(s); j = np.random.choice(np.arange(64), p=p)
        cdf = np.cumsum(p.astype(np.float64)); cdf /= cdf[-1]
        assert j == int(np.searchsorted(cdf, u, side="right")), s


def run_case(name, cfg_name, steps=3, eval_mode=True, seed=0, std_floor=None, **over):
    cfg = synth.make_cfg(cfg_name, **over)
    torch.manual_seed(0)
    agent = TDMPC(cfg)
    sd = synth.make_told_weights(cfg, seed)
    agent.model.load_state_dict(sd)
    agent.std = cfg.min_std if std_floor is None else std_floor
    step = 10 ** 6
    out = {}
    prev_mean_oracle = None
    for s in range(steps):
        obs = synth.make_obs(cfg, seed * 100 + s)
        nz = synth.make_noise(cfg, seed * 100 + s, step)
        H, N, P, T = synth.plan_dims(cfg, step)
        inj = Injector()
        for t in range(H):
            inj.stdn_q.append(nz["pi"][t])
        for i in range(cfg.iterations):
            inj.randn_q.append(nz["cem"][i])
            inj.stdn_q.append(nz["term"][i])
        inj.randn_q.append(nz["out"])
        inj.u = nz["u"]
        ev = agent.estimate_value

        def ev_logged(z, actions, horizon, _ev=ev, _inj=inj):
            v, r = _ev(z, actions, horizon)
            _inj.value_log.append(v.squeeze(1).clone())
            return v, r

        agent.estimate_value = ev_logged
        with inj:
            a, metrics = agent.plan(obs, eval_mode=eval_mode, step=step, t0=(s == 0))
        agent.estimate_value = ev
        out[f"s{s}_action"] = a.numpy().copy()
        out[f"s{s}_prev_mean"] = agent._prev_mean.numpy().copy()
        out[f"s{s}_values"] = torch.stack(inj.value_log).numpy()
        out[f"s{s}_elite_idx"] = torch.stack(inj.topk_log).numpy().astype(np.int32)
        out[f"s{s}_metrics"] = np.array([metrics["external_reward_mean"], metrics["current_std"]], dtype=np.float64)

        # the oracle restatement must agree with the reference it restates (same machine => tight)
        a_o, m_o, mean_o, taps = O.plan_tdmpc(sd, cfg, obs, eval_mode=eval_mode, step=step, t0=(s == 0),
                                              prev_mean=prev_mean_oracle, std_floor=agent.std, noise=nz)
        prev_mean_oracle = mean_o
        dv = (torch.stack(taps["values"]) - torch.stack(inj.value_log)).abs().max().item()
        da = (a_o - a).abs().max().item()
        dm = (mean_o - agent._prev_mean).abs().max().item()
        print(f"[{name}] step {s}: |dvalue|={dv:.3e} |daction|={da:.3e} |dmean|={dm:.3e} "
              f"metrics ref={metrics} oracle={m_o}")
        assert dv < 1e-5 and da < 1e-5 and dm < 1e-5
    out["meta"] = np.array([seed, steps, int(eval_mode)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def run_sim_case(name, cfg_name, plan, seed=0, eval_mode=True, **over):
    """TDMPCSIM.plan (tdmpc_similarity.py:135-200), the unmodified reference class -- the agent the quadrotor scripts
    build.  `plan`: list of (step, t0) calls; chosen so that the horizon schedule moves between calls (plan_horizon must
    follow it only at t0) and the policy-trajectory count follows regularization_schedule."""
    from src.algorithm.tdmpc_similarity import TDMPCSIM
    import src.algorithm.helper as ref_h2
    cfg = synth.make_cfg(cfg_name, batch_size=8, **over)
    torch.manual_seed(0)
    agent = TDMPCSIM(cfg)
    agent.device = torch.device("cpu")
    sd = synth.make_told_weights(cfg, seed, enc_ln=bool(cfg.normalize))
    missing = agent.model.load_state_dict(sd, strict=False)
    assert all(k.startswith("_predictor") for k in missing.missing_keys) and not missing.unexpected_keys
    agent.std = 0.1
    out = {}
    state = {"plan_horizon": 1, "std": agent.std}
    for s, (step, t0) in enumerate(plan):
        obs = synth.make_obs(cfg, seed * 100 + s)
        horizon = int(min(cfg.horizon, O.linear_schedule(cfg.horizon_schedule, step)))
        H = horizon if t0 else agent.plan_horizon
        P = int(O.linear_schedule(cfg.regularization_schedule, step) * cfg.num_samples)
        nz = synth.make_noise_hp(cfg, seed * 100 + s, H, P)
        inj = Injector()
        for t in range(H):
            if P > 0:
                inj.stdn_q.append(nz["pi"][t])
        for i in range(cfg.iterations):
            inj.randn_q.append(nz["cem"][i])
            inj.stdn_q.append(nz["term"][i])
        inj.randn_q.append(nz["out"])
        inj.u = nz["u"]
        ev = agent.estimate_value

        def ev_logged(z, actions, horizon, _ev=ev, _inj=inj):
            v = _ev(z, actions, horizon)
            _inj.value_log.append(v.squeeze(1).clone())
            return v

        agent.estimate_value = ev_logged
        ref_std = ref_h2._standard_normal
        with inj:
            ref_h2._standard_normal = ref_h._standard_normal      # same patched queue for both import names
            a = agent.plan(obs, eval_mode=eval_mode, step=step, t0=t0)
        ref_h2._standard_normal = ref_std
        agent.estimate_value = ev
        assert isinstance(a, torch.Tensor) and not inj.stdn_q and (eval_mode or not inj.randn_q)
        out[f"s{s}_action"] = a.numpy().copy()
        out[f"s{s}_prev_mean"] = agent._prev_mean.numpy().copy()
        out[f"s{s}_values"] = torch.stack(inj.value_log).numpy()
        out[f"s{s}_elite_idx"] = torch.stack(inj.topk_log).numpy().astype(np.int32)
        out[f"s{s}_hp"] = np.array([agent.plan_horizon, P], dtype=np.int64)
        a_o, taps = O.plan_tdmpcsim(sd, cfg, obs, eval_mode=eval_mode, step=step, t0=t0, state=state, noise=nz)
        dv = (torch.stack(taps["values"]) - torch.stack(inj.value_log)).abs().max().item()
        da = (a_o - a).abs().max().item()
        dm = (state["_prev_mean"] - agent._prev_mean).abs().max().item()
        print(f"[{name}] call {s} (step {step}, t0 {t0}): H={agent.plan_horizon} P={P} |dvalue|={dv:.3e} |daction|={da:.3e} |dmean|={dm:.3e}")
        assert dv < 1e-5 and da < 1e-5 and dm < 1e-5 and state["plan_horizon"] == agent.plan_horizon
    out["plan"] = np.array([[st, int(t)] for st, t in plan], dtype=np.int64)
    out["meta"] = np.array([seed, len(plan), int(eval_mode)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def run_update_case(name, cfg_name, batch=96, seed=0, **over):
    """TDMPC.update (tdmpc.py:191-245), the unmodified reference method, on CPU for one fixed replay batch: the returned
    metrics (loss means, weighted loss, grad norm) and the per-sample priorities it hands to update_priorities.  The online and
    the target model get DIFFERENT synthetic parameters so that a mix-up of the two (the TD target uses the online encoder and
    policy with the target Q heads) shows."""
    cfg = synth.make_cfg(cfg_name, batch_size=batch, **over)
    torch.manual_seed(0)
    agent = TDMPC(cfg)
    sd, sd_t = synth.make_told_weights(cfg, seed), synth.make_told_weights(cfg, seed + 50)
    agent.model.load_state_dict(sd)
    agent.model_target.load_state_dict(sd_t)
    b = synth.make_update_batch(cfg, seed, batch)

    class Buf:
        def sample(self):
            return b["obs"], b["next_obses"], b["action"], b["reward"], b["idxs"], b["weights"]

        def update_priorities(self, idxs, pr):
            self.idxs, self.pr = idxs.clone(), pr.clone()

    buf = Buf()
    q = [b["noise_pi"][t] for t in range(cfg.horizon)]
    real = ref_h._standard_normal

    def stdn(shape, dtype=None, device=None):
        if q:
            x = q.pop(0)
            assert tuple(x.shape) == tuple(shape), (x.shape, shape)
            return x.clone()
        return real(shape, dtype=dtype, device=device)                # the draws of update_pi (after the optimiser step)

    ref_h._standard_normal = stdn
    try:
        m = agent.update(buf, 10 ** 6)
    finally:
        ref_h._standard_normal = real
    assert not q and torch.equal(buf.idxs, b["idxs"])
    o = O.update_forward(sd, sd_t, cfg, b)
    out = {"priorities": buf.pr.numpy().copy(), "meta": np.array([seed, batch], dtype=np.int64),
           "metrics": np.array([m["consistency_loss"], m["reward_loss"], m["value_loss"], m["total_loss"], m["weighted_loss"], m["grad_norm"], m["pi_loss"]], dtype=np.float64)}
    mo = [float(o["consistency"].mean()), float(o["reward"].mean()), float(o["value"].mean()), float(o["total"].mean()), o["weighted_loss"]]
    dp = (o["priority"].clamp(max=1e4) - buf.pr).abs().max().item()
    print(f"[{name}] reference metrics {out['metrics'][:5]} oracle {mo} |dpriority| {dp:.3e}")
    assert np.allclose(out["metrics"][:5], mo, rtol=1e-6, atol=1e-7) and dp < 1e-6
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def run_train_case(name, cfg_name, batch=64, seed=0, steps=2, store_params=True, **over):
    """The WHOLE TDMPC.update (tdmpc.py:191-245: backward, clip_grad_norm_, both Adam optimisers, update_pi, ema), unmodified, on
    CPU for `steps` consecutive replay batches (step numbers chosen so that the target network moves on the first and not on
    the second call, update_freq = 2).  Stored: the returned metrics per call and -- for the small model -- every parameter of
    model and model_target after the last call; for the big ones a fingerprint (sum, sum of squares, 16 strided samples per
    tensor).  oracle.update_step must reproduce all of it."""
    cfg = synth.make_cfg(cfg_name, batch_size=batch, **over)
    torch.manual_seed(0)
    agent = TDMPC(cfg)
    sd, sd_t = synth.make_told_weights(cfg, seed), synth.make_told_weights(cfg, seed + 50)
    agent.model.load_state_dict(sd)
    agent.model_target.load_state_dict(sd_t)
    real = ref_h._standard_normal
    mets = []
    o_sd, o_t, o_opt = sd, sd_t, None
    worst = 0.0
    for s in range(steps):
        b = synth.make_train_batch(cfg, seed + 7 * s, batch)

        class Buf:
            def sample(self):
                return b["obs"], b["next_obses"], b["action"], b["reward"], b["idxs"], b["weights"]

            def update_priorities(self, idxs, pr):
                self.pr = pr.clone()

        buf = Buf()
        q = [b["noise_pi"][t] for t in range(cfg.horizon)] + [b["noise_pi_update"][t] for t in range(cfg.horizon + 1)]

        def stdn(shape, dtype=None, device=None, _q=q):
            x = _q.pop(0)
            assert tuple(x.shape) == tuple(shape), (x.shape, shape)
            return x.clone()

        ref_h._standard_normal = stdn
        try:
            m = agent.update(buf, 10 ** 6 + s)                 # even step: ema; odd step: no ema
        finally:
            ref_h._standard_normal = real
        assert not q
        mets.append([m["consistency_loss"], m["reward_loss"], m["value_loss"], m["total_loss"], m["weighted_loss"], m["grad_norm"], m["pi_loss"]])
        om, o_sd, o_t, o_g, o_opt = O.update_step(o_sd, o_t, cfg, b, o_opt, update_target=(10 ** 6 + s) % cfg.update_freq == 0)
        got = [om[k] for k in ("consistency_loss", "reward_loss", "value_loss", "total_loss", "weighted_loss", "grad_norm", "pi_loss")]
        dm = max(abs(a - c) / max(abs(a), 1e-12) for a, c in zip(mets[-1], got))
        dpar = max((agent.model.state_dict()[k] - o_sd[k]).abs().max().item() for k in o_sd)
        dtar = max((agent.model_target.state_dict()[k] - o_t[k]).abs().max().item() for k in o_t)
        dgr = max((p.grad - o_g[k]).abs().max().item() for k, p in agent.model.named_parameters())
        dpri = (buf.pr - om["priorities"]).abs().max().item()
        print(f"[{name}] call {s}: metrics rel diff {dm:.2e}  |d params| {dpar:.2e}  |d target| {dtar:.2e}  |d grads| {dgr:.2e}  |d priorities| {dpri:.2e}")
        worst = max(worst, dm, dpar, dtar, dgr)
    assert worst < 1e-6, worst
    out = {"metrics": np.array(mets, dtype=np.float64), "meta": np.array([seed, batch, steps], dtype=np.int64)}

    def finger(t):
        f = t.detach().double().flatten()
        idx = torch.linspace(0, f.numel() - 1, 16).long()
        return np.concatenate([[f.sum().item(), (f * f).sum().item()], f[idx].numpy()])

    for tag, mod in (("model", agent.model), ("target", agent.model_target)):
        for k, v in mod.state_dict().items():
            out[f"{tag}/{k}"] = v.numpy().copy() if store_params else finger(v)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


SIM_PLAN = [(9000, True), (9001, False), (17000, False), (17001, True), (17002, False)]     # horizon schedule: 3, 3, (4 ignored), 4, 4


def run_icem_case(name, cfg_name, steps=3, seed=0):
    """TdICemSimMlp.plan (tdmpc_icem_similarity_mlp.py:159-265), the unmodified reference class."""
    import colorednoise
    from src.algorithm.tdmpc_icem_similarity_mlp import TdICemSimMlp
    import src.algorithm.helper as ref_h2          # the iCEM file imports helper under this name
    cfg = synth.make_cfg(cfg_name)
    torch.manual_seed(0)
    agent = TdICemSimMlp(cfg)
    agent.device = torch.device("cpu")
    sd = synth.make_told_weights(cfg, seed, enc_ln=bool(cfg.normalize))
    missing = agent.model.load_state_dict(sd, strict=False)
    assert all(k.startswith("_predictor") for k in missing.missing_keys) and not missing.unexpected_keys
    agent.std = cfg.min_std
    step = 10 ** 6
    out = {}
    state = {"plan_horizon": 1, "std": agent.std}
    for s in range(steps):
        obs = synth.make_obs(cfg, seed * 100 + s)
        t0 = s == 0
        horizon = int(min(cfg.horizon, O.linear_schedule(cfg.horizon_schedule, step)))
        H = horizon if t0 else agent.plan_horizon
        mix = O.linear_schedule(cfg.regularization_schedule, step)
        have = hasattr(agent, "_elite_actions")
        sizes = O.icem_sizes(cfg, mix, have)
        nz = synth.make_icem_noise(cfg, seed * 100 + s, sizes, H, have)
        inj = Injector()
        for t in range(H):
            inj.stdn_q.append(nz["pi"][t])
        colored_q = []
        for i in range(cfg.iterations):
            inj.randn_q.append(nz["white"][i])
            colored_q += [nz["pink"][i], nz["brown"][i]]
            if i == 0 and "shift_raw" in nz:
                colored_q.append(nz["shift_raw"])
            inj.stdn_q.append(nz["term"][i])
        inj.randn_q.append(nz["out"])
        inj.u = nz["u"]

        def fake_powerlaw(exponent, size, **kw):
            x = colored_q.pop(0)
            assert tuple(x.shape) == tuple(size), (x.shape, size)
            return x.copy()

        real_pl = colorednoise.powerlaw_psd_gaussian
        colorednoise.powerlaw_psd_gaussian = fake_powerlaw
        ev = agent.estimate_value

        def ev_logged(z, actions, _ev=ev, _inj=inj):
            v, r = _ev(z, actions)
            _inj.value_log.append(v.squeeze(1).clone())
            return v, r

        agent.estimate_value = ev_logged
        ref_std = ref_h2._standard_normal
        with inj:
            ref_h2._standard_normal = ref_h._standard_normal      # same patched queue for both import names
            a, metrics = agent.plan(obs, eval_mode=True, step=step, t0=t0)
        ref_h2._standard_normal = ref_std
        agent.estimate_value = ev
        colorednoise.powerlaw_psd_gaussian = real_pl
        assert not colored_q and not inj.stdn_q
        Tm = max(v.shape[0] for v in inj.value_log)
        vals = np.zeros((cfg.iterations, Tm), dtype=np.float32)
        for i, v in enumerate(inj.value_log):
            vals[i, :v.shape[0]] = v.numpy()
        out[f"s{s}_action"] = a.numpy().copy()
        out[f"s{s}_prev_mean"] = agent._prev_mean.numpy().copy()
        out[f"s{s}_values"] = vals
        out[f"s{s}_elite_idx"] = torch.stack(inj.topk_log).numpy().astype(np.int32)
        out[f"s{s}_elite_actions"] = agent._elite_actions.numpy().copy()
        out[f"s{s}_sizes"] = np.array(sizes, dtype=np.int32)
        out[f"s{s}_metrics"] = np.array([metrics["external_reward_mean"], metrics["current_std"]], dtype=np.float64)

        a_o, m_o, taps = O.plan_icem_mlp(sd, cfg, obs, eval_mode=True, step=step, t0=t0, state=state, noise=nz)
        dv = max((taps["values"][i] - inj.value_log[i]).abs().max().item() for i in range(cfg.iterations))
        da = (a_o - a).abs().max().item()
        dm = (state["_prev_mean"] - agent._prev_mean).abs().max().item()
        de = (state["_elite_actions"] - agent._elite_actions).abs().max().item()
        print(f"[{name}] step {s}: sizes {sizes[0]}..{sizes[-1]} |dvalue|={dv:.3e} |daction|={da:.3e} |dmean|={dm:.3e} |delite|={de:.3e}")
        assert dv < 1e-5 and da < 1e-5 and dm < 1e-5 and de < 1e-5
    out["meta"] = np.array([seed, steps, 1], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def run_drnn_case(name, cfg_name, steps=3, seed=0):
    """TdICemSimDssm.plan (tdmpc_icem_similarity_drnn.py:168-272), the unmodified reference class
    (NormGRUCell + BatchNorm prior MLP dynamics, reward from the hidden state)."""
    import colorednoise
    from src.algorithm.tdmpc_icem_similarity_drnn import TdICemSimDssm
    import src.algorithm.helper as ref_h2
    cfg = synth.make_cfg(cfg_name)
    torch.manual_seed(0)
    agent = TdICemSimDssm(cfg)
    agent.device = torch.device("cpu")
    sd = synth.make_drnn_weights(cfg, seed)
    missing = agent.model.load_state_dict(sd, strict=False)
    assert all(k.startswith("_predictor") or k.endswith("num_batches_tracked") for k in missing.missing_keys) and not missing.unexpected_keys, missing
    agent.model.eval()
    agent.std = cfg.min_std
    step = 10 ** 6
    out = {}
    state = {"plan_horizon": 1, "std": agent.std}
    hidden = torch.zeros(1, cfg.hidden_dim)
    hidden_o = hidden.clone()
    for s in range(steps):
        obs = synth.make_obs(cfg, seed * 100 + s)
        t0 = s == 0
        horizon = int(min(cfg.horizon, O.linear_schedule(cfg.horizon_schedule, step)))
        H = horizon if t0 else agent.plan_horizon
        mix = O.linear_schedule(cfg.regularization_schedule, step)
        have = hasattr(agent, "_elite_actions")
        sizes = O.icem_drnn_sizes(cfg, mix, have, t0)
        nz = synth.make_icem_noise(cfg, seed * 100 + s, sizes, H, have and not t0, single_beta=True)
        inj = Injector()
        for t in range(H):
            inj.stdn_q.append(nz["pi"][t])
        colored_q = []
        for i in range(cfg.iterations):
            colored_q.append(nz["pink"][i])
            if i == 0 and "shift_raw" in nz:
                colored_q.append(nz["shift_raw"])
            inj.stdn_q.append(nz["term"][i])
        inj.randn_q.append(nz["out"])
        inj.u = nz["u"]

        def fake_powerlaw(exponent, size, **kw):
            x = colored_q.pop(0)
            assert tuple(x.shape) == tuple(size), (x.shape, size)
            return x.copy()

        real_pl = colorednoise.powerlaw_psd_gaussian
        colorednoise.powerlaw_psd_gaussian = fake_powerlaw
        ev = agent.estimate_value

        def ev_logged(z, actions, hid, _ev=ev, _inj=inj):
            v, r = _ev(z, actions, hid)
            _inj.value_log.append(v.squeeze(1).clone())
            return v, r

        agent.estimate_value = ev_logged
        ref_std = ref_h2._standard_normal
        with inj:
            ref_h2._standard_normal = ref_h._standard_normal
            a, hidden, metrics = agent.plan(obs, hidden, eval_mode=True, step=step, t0=t0)
        ref_h2._standard_normal = ref_std
        agent.estimate_value = ev
        colorednoise.powerlaw_psd_gaussian = real_pl
        assert not colored_q and not inj.stdn_q
        Tm = max(v.shape[0] for v in inj.value_log)
        vals = np.zeros((cfg.iterations, Tm), dtype=np.float32)
        for i, v in enumerate(inj.value_log):
            vals[i, :v.shape[0]] = v.numpy()
        out[f"s{s}_action"] = a.numpy().copy()
        out[f"s{s}_hidden"] = hidden.numpy().copy()
        out[f"s{s}_prev_mean"] = agent._prev_mean.numpy().copy()
        out[f"s{s}_values"] = vals
        out[f"s{s}_elite_idx"] = torch.stack(inj.topk_log).numpy().astype(np.int32)
        out[f"s{s}_elite_actions"] = agent._elite_actions.numpy().copy()
        out[f"s{s}_sizes"] = np.array(sizes, dtype=np.int32)
        out[f"s{s}_metrics"] = np.array([metrics["external_reward_mean"], metrics["current_std"]], dtype=np.float64)

        a_o, hidden_o, m_o, taps = O.plan_icem_drnn(sd, cfg, obs, hidden_o, eval_mode=True, step=step, t0=t0, state=state, noise=nz)
        dv = max((taps["values"][i] - inj.value_log[i]).abs().max().item() for i in range(cfg.iterations))
        da = (a_o - a).abs().max().item()
        dh = (hidden_o - hidden).abs().max().item()
        dm = (state["_prev_mean"] - agent._prev_mean).abs().max().item()
        print(f"[{name}] step {s}: sizes {sizes[0]}..{sizes[-1]} |dvalue|={dv:.3e} |daction|={da:.3e} |dhidden|={dh:.3e} |dmean|={dm:.3e}")
        assert dv < 1e-5 and da < 1e-5 and dm < 1e-5 and dh < 1e-5
    out["meta"] = np.array([seed, steps, 1], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def run_replay_case(name, cfg_name, seed=0, draws=3):
    """ReplayBuffer.sample of the reference (helper.py:506-534) on CPU: the buffer is filled through the reference's own
    add / update_priorities with tests/synth.py data, np.random is seeded so that the uniforms it consumes are
    synth.make_uniforms(seed).  Stored: the reference's outputs and its priority vector."""
    from oracle import replay_oracle as RO
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self          # helper.py:530-532 moves the batch with .cuda()
    cfg = synth.make_replay_cfg(cfg_name)
    buf = ref_h.ReplayBuffer(cfg, latent_plan=True)
    st = RO.ReplayState(cfg, latent_plan=True)
    for obs, act, rew in synth.make_replay_episodes(cfg, seed):
        ep = ref_h.Episode(cfg, obs[0])
        for t in range(cfg.episode_length):
            ep += (obs[t + 1], torch.from_numpy(act[t]), float(rew[t]), t == cfg.episode_length - 1)
        buf += ep
        st.add(torch.from_numpy(obs), torch.from_numpy(act), torch.from_numpy(rew))
    n_valid = buf.capacity if buf._full else buf.idx
    for idxs, pr in synth.make_priority_updates(cfg, n_valid, seed):
        buf.update_priorities(torch.from_numpy(idxs), torch.from_numpy(pr))
        st.update_priorities(torch.from_numpy(idxs), torch.from_numpy(pr))
    assert torch.equal(buf._priorities, st["priorities"]) and torch.equal(buf._obs[:n_valid], st["obs"][:n_valid]) and buf.idx == st["idx"]
    buf._obs[n_valid:] = 0          # torch.empty rows the reference never wrote (idx + t + 1 can touch row n_valid)
    out = {"priorities": buf._priorities.numpy().copy(), "idx_full": np.array([buf.idx, int(buf._full)], dtype=np.int64)}
    np.random.seed(seed)
    uni = synth.make_uniforms(cfg.batch_size * 16 * draws, seed)
    used = 0
    for d in range(draws):
        ref = buf.sample()
        orc = RO.sample(st, uni[used:])
        used += orc[6]
        for k, (r, o) in enumerate(zip(ref, orc[:6])):
            assert torch.equal(r, o) if k == 4 else torch.allclose(r, o, rtol=1e-6, atol=0), (name, d, k)
        for key, v in zip(("obs", "next_obs", "action", "reward", "idxs", "weights"), ref):
            out[f"d{d}_{key}"] = v.numpy().copy()
        out[f"d{d}_used"] = np.array([orc[6]], dtype=np.int64)
        print(f"[{name}] draw {d}: full={buf._full} uniforms used {orc[6]} (batch {cfg.batch_size}), "
              f"distinct idx {len(set(ref[4].tolist()))}, oracle == reference")
    assert abs(np.random.random_sample() - uni[used]) == 0.0        # the reference consumed exactly the same stream
    out["meta"] = np.array([seed, draws], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "replay":
        run_replay_case("replay_small", "small", seed=0)
        run_replay_case("replay_full", "full", seed=1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "update":
        run_update_case("update_tiny", "tiny", batch=70, seed=12)
        run_update_case("update_cfg3", "cfg3", batch=96, seed=13)
        run_update_case("update_cfg2s", "cfg2s", batch=40, seed=14)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        run_train_case("train_small", "tiny", batch=48, seed=21, mlp_dim=64, enc_dim=32)
        run_train_case("train_cfg3", "cfg3", batch=64, seed=22, store_params=False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sim":
        run_sim_case("tdmpcsim_tiny", "tiny", SIM_PLAN, seed=8, horizon=5, horizon_schedule="linear(2, 5, 25000, 0)", normalize=True)
        run_sim_case("tdmpcsim_cfg3", "cfg3", SIM_PLAN[:4], seed=9, eval_mode=False)
        sys.exit(0)
    check_choice_replica()
    run_case("tdmpc_tiny", "tiny", steps=3)
    run_case("tdmpc_tiny_train", "tiny", steps=2, eval_mode=False, seed=3, std_floor=0.1)
    run_case("tdmpc_cfg1", "cfg1", steps=3)
    run_case("tdmpc_cfg3", "cfg3", steps=2, seed=1)
    run_case("tdmpc_cfg2s", "cfg2s", steps=2, seed=2)
    run_icem_case("icem_tiny", "tiny_icem", steps=3, seed=4)
    run_icem_case("icem_cfg4", "cfg4", steps=2, seed=5)
    run_drnn_case("drnn_tiny", "tiny_drnn", steps=3, seed=6)
    run_drnn_case("drnn_cfg5", "cfg5", steps=2, seed=7)
    run_replay_case("replay_small", "small", seed=0)
    run_replay_case("replay_full", "full", seed=1)
    run_sim_case("tdmpcsim_tiny", "tiny", SIM_PLAN, seed=8, horizon=5, horizon_schedule="linear(2, 5, 25000, 0)", normalize=True)
    run_sim_case("tdmpcsim_cfg3", "cfg3", SIM_PLAN[:4], seed=9, eval_mode=False)
    run_update_case("update_tiny", "tiny", batch=70, seed=12)
    run_update_case("update_cfg3", "cfg3", batch=96, seed=13)
    run_update_case("update_cfg2s", "cfg2s", batch=40, seed=14)
    run_train_case("train_small", "tiny", batch=48, seed=21, mlp_dim=64, enc_dim=32)
    run_train_case("train_cfg3", "cfg3", batch=64, seed=22, store_params=False)
    print("golden fixtures written to", HERE)

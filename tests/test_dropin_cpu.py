"""Host logic of the drop-in layer (tdmpc_b200/planner.py) on CPU: `accelerate()` class dispatch for the reference's four
agent shapes, return arity, seed-step early returns, `_prev_mean` / `_elite_actions` / `plan_horizon` side effects and
the episode-start horizon growth that the reference survives (tdmpc.py:124 only reads `_prev_mean` when `not t0`).
The CUDA planner is replaced by a recorder -- nothing here computes a plan; parity lives in the -m gpu tests."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tdmpc_b200 import planner as P
from tdmpc_b200.told import DSSM, DSSMRnn, TOLD
from tests import synth


class _Recorder:
    """Stands in for planner.Planner: records constructor / plan arguments, returns zero tensors of the right shape."""
    instances = []

    def __init__(self, obs_dim, action_dim, latent_dim, mlp_dim=512, enc_dim=256, enc_layernorm=False, precision="bf16",
                 device=None, dynamics="mlp", hidden_dim=0):
        self.ctor = dict(obs_dim=obs_dim, action_dim=action_dim, enc_layernorm=enc_layernorm, precision=precision,
                         dynamics=dynamics, hidden_dim=hidden_dim)
        self.A, self.calls, self.seed, self.call = action_dim, [], 0, 0
        _Recorder.instances.append(self)

    def sync_weights(self, model, force=False):
        return False

    def plan(self, obs, **kw):
        self.calls.append(dict(kw, obs_shape=tuple(obs.shape)))
        H = kw["horizon"]
        pm = kw["prev_mean"] if kw.get("prev_mean") is not None else torch.zeros(1, H, self.A)
        if kw.get("elite_actions") is not None:
            kw["elite_actions"].fill_(0.25)
        return {"action": torch.zeros(1, self.A), "metrics": torch.tensor([[1.5, 0.25]]), "prev_mean": pm + 1.0,
                "taps": None, "hidden": kw.get("hidden_out")}


@pytest.fixture
def rec(monkeypatch):
    _Recorder.instances.clear()
    monkeypatch.setattr(P, "Planner", _Recorder)
    return _Recorder


def _agent(kind, **over):
    """Attribute-compatible stand-ins for the reference agents (same attributes the reference constructors set)."""
    name = {"tdmpc": "tiny", "sim": "tiny", "icem": "tiny_icem", "drnn": "tiny_drnn"}[kind]
    cfg = synth.make_cfg(name, **over)
    model = {"tdmpc": TOLD, "sim": DSSM, "icem": DSSM, "drnn": DSSMRnn}[kind](cfg)
    cls = {"tdmpc": type("TDMPC", (), {}), "sim": type("TDMPCSIM", (), {}),
           "icem": type("TdICemSimMlp", (), {"sample_mix_action_sequence": lambda self, *a: None}),
           "drnn": type("TdICemSimDssm", (), {"sample_mix_action_sequence": lambda self, *a: None})}[kind]
    ag = cls()
    ag.cfg, ag.model, ag.device, ag.std = cfg, model, torch.device("cpu"), 0.1
    if kind != "tdmpc":
        ag.plan_horizon, ag.mixture_coef = 1, 0.05
    return ag


@pytest.mark.parametrize("kind", ["tdmpc", "sim", "icem", "drnn"])
def test_accelerate_dispatch_and_return_arity(rec, kind):
    over = dict(normalize=True, horizon=5, horizon_schedule="linear(2, 5, 25000, 0)") if kind == "sim" else {}
    ag = P.accelerate(_agent(kind, **over), precision="fp32")
    assert P.agent_kind(ag) == kind and ag.plan.__func__ is P._PLAN_KINDS[kind]
    pl = rec.instances[-1]
    assert pl.ctor["dynamics"] == ("drnn" if kind == "drnn" else "mlp") and pl.ctor["precision"] == "fp32"
    assert pl.ctor["enc_layernorm"] == (kind != "tdmpc")            # dmlab_enc_norm (LayerNorm) for the cfg.normalize agents
    assert pl.ctor["hidden_dim"] == (ag.cfg.hidden_dim if kind == "drnn" else 0)
    obs = synth.make_obs(ag.cfg, 0)
    cfg = ag.cfg
    step = 10 ** 6
    if kind == "drnn":
        hid = torch.zeros(1, cfg.hidden_dim)
        a, h2, m = ag.plan(obs, hid, eval_mode=False, step=0, t0=True)               # seed step (:170-172)
        assert a.shape == (cfg.action_dim,) and h2 is None and m == {"external_reward_mean": 0.0, "current_std": 0.0}
        a, h2, m = ag.plan(obs, hid, eval_mode=True, step=step, t0=True)
        assert a.shape == (cfg.action_dim,) and tuple(h2.shape) == (1, cfg.hidden_dim) and m["external_reward_mean"] == 1.5
        assert ag.plan_horizon == cfg.horizon and tuple(ag._elite_actions.shape) == (cfg.horizon, cfg.num_elites, cfg.action_dim)
    elif kind == "sim":
        a = ag.plan(obs, eval_mode=False, step=0, t0=True)                            # bare tensor, seed step (:144-145)
        assert isinstance(a, torch.Tensor) and a.shape == (cfg.action_dim,) and not pl.calls
        a = ag.plan(obs, eval_mode=False, step=0, t0=True, fine_tuning=True)          # fine_tuning plans even during seed steps
        assert isinstance(a, torch.Tensor) and len(pl.calls) == 1
        assert ag.plan_horizon == 2 and pl.calls[-1]["horizon"] == 2                  # horizon_schedule at step 0
        assert pl.calls[-1]["num_pi_trajs"] == int(0.05 * cfg.num_samples)            # regularization_schedule at step 0
        a = ag.plan(obs, eval_mode=True, step=step, t0=False)
        assert pl.calls[-1]["horizon"] == 2 and ag.plan_horizon == 2                  # only moves at an episode start (:151-152)
        assert pl.calls[-1]["num_pi_trajs"] == int(0.5 * cfg.num_samples) and pl.calls[-1]["has_prev"] is True
        a = ag.plan(obs, eval_mode=True, step=step, t0=True)
        assert pl.calls[-1]["horizon"] == cfg.horizon and pl.calls[-1]["has_prev"] is False
    else:
        a, m = ag.plan(obs, eval_mode=False, step=0, t0=True)
        assert a.shape == (cfg.action_dim,) and m == {"external_reward_mean": 0.0, "current_std": 0.0} and not pl.calls
        a, m = ag.plan(obs, eval_mode=True, step=step, t0=True)
        assert a.shape == (cfg.action_dim,) and m == {"external_reward_mean": 1.5, "current_std": 0.25}
        assert tuple(ag._prev_mean.shape) == (cfg.horizon, cfg.action_dim)
        if kind == "icem":
            assert ag.plan_horizon == cfg.horizon and tuple(ag._elite_actions.shape) == (cfg.horizon, cfg.num_elites, cfg.action_dim)
            assert pl.calls[-1]["warm_last"] and pl.calls[-1]["best_a_last"] and pl.calls[-1]["init_std"] == 0.5


def test_tdmpc_horizon_growth_at_episode_start(rec):
    """with horizon_schedule linear(2,6,25000,0) the horizon grows 3 -> 4 at step 12500, always at a
    t0=True call (train.py passes episode.first).  The reference ignores the stale 3-step `_prev_mean` there
    (tdmpc.py:124); the drop-in must not shape-check it."""
    ag = P.accelerate(_agent("tdmpc", horizon=6, horizon_schedule="linear(2, 6, 25000, 0)", seed_steps=0), precision="fp32")
    pl = rec.instances[-1]
    obs = synth.make_obs(ag.cfg, 0)
    ag.plan(obs, eval_mode=True, step=12000, t0=True)
    assert pl.calls[-1]["horizon"] == 3 and pl.calls[-1]["has_prev"] is False
    ag.plan(obs, eval_mode=True, step=12001, t0=False)
    assert pl.calls[-1]["horizon"] == 3 and pl.calls[-1]["has_prev"] is True and tuple(pl.calls[-1]["prev_mean"].shape) == (1, 3, ag.cfg.action_dim)
    ag.plan(obs, eval_mode=True, step=12500, t0=True)                  # 3 -> 4 at an episode start: must not raise
    assert pl.calls[-1]["horizon"] == 4 and pl.calls[-1]["has_prev"] is False and pl.calls[-1]["prev_mean"] is None
    assert tuple(ag._prev_mean.shape) == (4, ag.cfg.action_dim)
    ag.plan(obs, eval_mode=True, step=12501, t0=False)
    assert pl.calls[-1]["has_prev"] is True and tuple(pl.calls[-1]["prev_mean"].shape) == (1, 4, ag.cfg.action_dim)
    # mid-episode growth is a shape error in the reference too (mean[:-1] = prev[1:] cannot broadcast 2 rows into 4)
    ag._prev_mean = torch.zeros(3, ag.cfg.action_dim)
    with pytest.raises(RuntimeError):
        ag.plan(obs, eval_mode=True, step=20000, t0=False)


def test_accelerate_rejects_non_layernorm_encoder(rec):
    ag = _agent("sim", normalize=True)
    ag.model._encoder[1] = torch.nn.BatchNorm1d(ag.cfg.enc_dim)       # cfg.norm_type='bn' in helper.dmlab_enc_norm
    with pytest.raises(NotImplementedError):
        P.accelerate(ag)


def test_icem_requires_keep_previous_elites(rec):
    ag = P.accelerate(_agent("icem", keep_previous_elites=False), precision="fp32")
    with pytest.raises(NotImplementedError):
        ag.plan(synth.make_obs(ag.cfg, 0), eval_mode=True, step=10 ** 6, t0=True)

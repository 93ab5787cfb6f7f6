"""Host side of the B200 planner: a thin Python layer over the C ABI (include/tdmpc_b200.h).

`Planner`    -- batched, stage-level interface (encode / estimate_value / pi_rollout /
                select_update / plan) on device tensors; PyTorch is used only for device memory and
                the current CUDA stream.
`TDMPC`      -- drop-in for the reference agent's planning surface (src/algorithm/tdmpc.py:53-163):
                same constructor argument, attributes (cfg, device, std, model, model_target,
                _prev_mean), `plan(obs, eval_mode, step, t0)` signature, return types and side
                effects.  Training (`update`) is out of scope and stays in the reference.
`accelerate` -- binds the fast `plan` onto an existing reference agent object in place.
"""
import ctypes as C
import types
from copy import deepcopy

import numpy as np
import torch

from . import _lib
from .colored import color_matrices, mix_split, synthesis_matrix
from .told import DSSM, DSSMRnn, TOLD, linear_schedule

_PREC = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "bf16_exact": _lib.PREC_BF16_EXACT}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class Planner:
    """One handle = one model on one GPU.  Not thread-safe; stream-ordered on torch's current stream."""

    def __init__(self, obs_dim, action_dim, latent_dim, mlp_dim=512, enc_dim=256, enc_layernorm=False,
                 precision="bf16", device=None, dynamics="mlp", hidden_dim=0):
        self.lib = _lib.load()          # raises if libtdmpc_b200.so is missing: there is no fallback
        if not torch.cuda.is_available():
            raise RuntimeError("tdmpc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.D, self.A, self.L, self.M, self.E = obs_dim, action_dim, latent_dim, mlp_dim, enc_dim
        self.precision = precision
        self.drnn = dynamics == "drnn"
        self.Hd = int(hidden_dim) if self.drnn else 0
        cfg = _lib.ModelCfg(obs_dim, action_dim, latent_dim, mlp_dim, enc_dim, self.Hd, int(bool(enc_layernorm)),
                            {"mlp": _lib.DYN_MLP, "drnn": _lib.DYN_DRNN}[dynamics], _PREC[precision], self.device.index)
        self._cfg = cfg
        self._h = C.c_void_p()
        rc = self.lib.tdmpc_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"tdmpc_create failed ({rc}): {self.lib.tdmpc_last_error(None).decode()}")
        self._keep = None               # tensors whose pointers the last set_weights used
        self._wver = None
        self._cap = (0, 0, 0, 0)
        self.call = 0
        self.seed = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self.lib.tdmpc_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.tdmpc_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self):
        return int(self.lib.tdmpc_launch_count(self._h))

    def reserve(self, batch, rows, horizon, elites):
        cap = self._cap
        if batch > cap[0] or rows > cap[1] or horizon > cap[2] or elites > cap[3]:
            new = (max(batch, cap[0]), max(rows, cap[1]), max(horizon, cap[2]), max(elites, cap[3]))
            self._check(self.lib.tdmpc_reserve(self._h, *new), "tdmpc_reserve")
            self._cap = new

    # ------------------------------------------------------------------ weights
    def sync_weights(self, model_or_sd, force=False):
        """(Re)pack the parameters of a TOLD-like module / state_dict.  Cheap no-op when no
        parameter changed since the last call (tracked through Tensor._version + data_ptr)."""
        sd = model_or_sd if isinstance(model_or_sd, dict) else dict(model_or_sd.state_dict(keep_vars=True))
        ln = "_encoder.3.weight" in sd
        last = "_encoder.3" if ln else "_encoder.2"
        mlps = ("_reward", "_pi") if self.drnn else ("_dynamics", "_reward", "_pi")
        names = ["_encoder.0", last] + [f"{m}.{i}" for m in mlps for i in (0, 2, 4)] + \
                [f"{m}.{i}" for m in ("_Q1", "_Q2") for i in (0, 3, 6)]
        keys = [n + s for n in names for s in (".weight", ".bias")] + \
               [f"{m}.{i}{s}" for m in ("_Q1", "_Q2") for i in (1, 4) for s in (".weight", ".bias")] + \
               (["_encoder.1.weight", "_encoder.1.bias"] if ln else [])
        if self.drnn:       # models/gru_dyna.py:11-23, models/rnns.py:8-17
            g = "_dynamics.gru_cell."
            pm = "_dynamics.prior_mlp."
            keys += [g + "weight_ih.weight", g + "weight_hh.weight"] + \
                    [f"{g}ln_{n}{s}" for n in ("reset", "update", "newval") for s in (".weight", ".bias")] + \
                    [pm + k for k in ("0.weight", "0.bias", "3.weight", "3.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var")]
        ver = tuple((sd[k].data_ptr(), sd[k]._version) for k in keys)
        if not force and ver == self._wver:
            return False
        t = {k: _f32(sd[k], self.device) for k in keys}
        w = _lib.Weights()

        def P(k):
            return t[k].data_ptr()

        w.enc_w[0], w.enc_b[0] = P("_encoder.0.weight"), P("_encoder.0.bias")
        w.enc_w[1], w.enc_b[1] = P(last + ".weight"), P(last + ".bias")
        if ln:
            w.enc_ln_g, w.enc_ln_b = P("_encoder.1.weight"), P("_encoder.1.bias")
        for dst_w, dst_b, mod in ((w.dyn_w, w.dyn_b, "_dynamics"), (w.rew_w, w.rew_b, "_reward"), (w.pi_w, w.pi_b, "_pi")):
            if mod == "_dynamics" and self.drnn:
                continue
            for j, i in enumerate((0, 2, 4)):
                dst_w[j], dst_b[j] = P(f"{mod}.{i}.weight"), P(f"{mod}.{i}.bias")
        if self.drnn:
            w.gru_w_ih, w.gru_w_hh = P(g + "weight_ih.weight"), P(g + "weight_hh.weight")
            for j, n in enumerate(("reset", "update", "newval")):
                w.gru_ln_g[j], w.gru_ln_b[j] = P(f"{g}ln_{n}.weight"), P(f"{g}ln_{n}.bias")
            w.prior_w[0], w.prior_b[0], w.prior_w[1], w.prior_b[1] = P(pm + "0.weight"), P(pm + "0.bias"), P(pm + "3.weight"), P(pm + "3.bias")
            for j, n in enumerate(("weight", "bias", "running_mean", "running_var")):
                w.prior_bn[j] = P(pm + "1." + n)
        for qi, mod in enumerate(("_Q1", "_Q2")):
            for j, i in enumerate((0, 3, 6)):
                w.q_w[qi][j], w.q_b[qi][j] = P(f"{mod}.{i}.weight"), P(f"{mod}.{i}.bias")
            for j, i in enumerate((1, 4)):
                w.q_ln_g[qi][j], w.q_ln_b[qi][j] = P(f"{mod}.{i}.weight"), P(f"{mod}.{i}.bias")
        shapes = {"_encoder.0.weight": (self.E, self.D), last + ".weight": (self.L, self.E),
                  "_pi.0.weight": (self.M, self.L), "_pi.4.weight": (self.A, self.M), "_Q1.6.weight": (1, self.M)}
        if self.drnn:
            shapes.update({g + "weight_ih.weight": (3 * self.Hd, self.L + self.A), g + "weight_hh.weight": (3 * self.Hd, self.Hd),
                           pm + "0.weight": (self.M, self.Hd), pm + "3.weight": (self.L, self.M), "_reward.0.weight": (self.M, self.Hd)})
        else:
            shapes.update({"_dynamics.0.weight": (self.M, self.L + self.A), "_dynamics.4.weight": (self.L, self.M)})
        for k, s in shapes.items():
            if tuple(t[k].shape) != s:
                raise RuntimeError(f"sync_weights: {k} has shape {tuple(t[k].shape)}, expected {s}")
        self._check(self.lib.tdmpc_set_weights(self._h, C.byref(w), self._stream()), "tdmpc_set_weights")
        self._keep, self._wver = t, ver
        return True

    # ------------------------------------------------------------------ stage-level entry points
    def encode(self, obs):
        """TOLD.h (tdmpc.py:30-32) for obs [B, D] -> z [B, L]."""
        obs = _f32(obs, self.device)
        z = torch.empty(obs.shape[0], self.L, device=self.device)
        self._check(self.lib.tdmpc_encode(self._h, _ptr(obs), obs.shape[0], _ptr(z), self._stream()), "tdmpc_encode")
        return z

    def estimate_value(self, z0, actions, *, discount, min_std, noise_term=None, seed=0, call=0, hidden=None):
        """TDMPC.estimate_value (tdmpc.py:83-92).  z0 [B,L]; actions [B,T,H,A]; noise_term [B,T,A]; GRU dynamics:
        hidden [B,Hd] (tdmpc_icem_similarity_drnn.py:124-133; None = zeros)."""
        z0, actions = _f32(z0, self.device), _f32(actions, self.device)
        B, T, H, _ = actions.shape
        values = torch.empty(B, T, device=self.device)
        last_r = torch.empty(B, T, device=self.device)
        nz = _f32(noise_term, self.device) if noise_term is not None else None
        if hidden is not None:
            hidden = _f32(hidden, self.device)
            assert tuple(hidden.shape) == (B, self.Hd)
            self._check(self.lib.tdmpc_estimate_value_drnn(self._h, _ptr(z0), _ptr(hidden), _ptr(actions), _ptr(nz), B, T, H, float(discount),
                                                           float(min_std), seed, call, _ptr(values), _ptr(last_r), self._stream()),
                        "tdmpc_estimate_value_drnn")
        else:
            self._check(self.lib.tdmpc_estimate_value(self._h, _ptr(z0), _ptr(actions), _ptr(nz), B, T, H, float(discount),
                                                      float(min_std), seed, call, _ptr(values), _ptr(last_r), self._stream()),
                        "tdmpc_estimate_value")
        return values, last_r

    def pi_rollout(self, z0, rows, horizon, *, min_std, noise_pi=None, seed=0, call=0):
        """Policy rollout of TDMPC.plan (tdmpc.py:113-118): returns pi_actions [B, P, H, A]."""
        z0 = _f32(z0, self.device)
        B = z0.shape[0]
        out = torch.empty(B, rows, horizon, self.A, device=self.device)
        nz = _f32(noise_pi, self.device) if noise_pi is not None else None
        self._check(self.lib.tdmpc_pi_rollout(self._h, _ptr(z0), _ptr(nz), B, rows, horizon, float(min_std), seed, call,
                                              _ptr(out), self._stream()), "tdmpc_pi_rollout")
        return out

    def select_update(self, values, actions, mean, std, *, num_elites, temperature, momentum, std_floor, std_max=2.0):
        """tdmpc.py:138-149.  Updates mean/std [B,H,A] in place; returns (elite_idx [B,K], score [B,K])."""
        values, actions = _f32(values, self.device), _f32(actions, self.device)
        B, T, H, _ = actions.shape
        idx = torch.empty(B, num_elites, dtype=torch.int32, device=self.device)
        score = torch.empty(B, num_elites, device=self.device)
        self._check(self.lib.tdmpc_select_update(self._h, _ptr(values), _ptr(actions), B, T, H, num_elites,
                                                 float(temperature), float(momentum), float(std_floor), float(std_max),
                                                 _ptr(mean), _ptr(std), _ptr(idx), _ptr(score), self._stream()),
                    "tdmpc_select_update")
        return idx, score

    def fill_noise(self, stream_id, count, seed, call):
        """The built-in Philox stream (so tests can hand the same draws to the oracle)."""
        out = torch.empty(count, dtype=torch.float64 if stream_id == 4 else torch.float32, device=self.device)
        self._check(self.lib.tdmpc_fill_noise(self._h, stream_id, seed, call, count, _ptr(out), self._stream()), "tdmpc_fill_noise")
        return out

    # ------------------------------------------------------------------ the hot path
    def plan(self, obs=None, *, z0=None, horizon, num_samples, num_pi_trajs, num_elites, iterations, discount,
             temperature, momentum, min_std, std_floor, std_max=2.0, init_std=2.0, eval_mode=False, t0=True,
             prev_mean=None, has_prev=None, noise=None, taps=False, seed=None, call=None,
             n_samples=None, n_reuse=None, n_pi=None, out=None, elite_actions=None, elite_horizon=0,
             warm_last=False, best_a_last=False, shift_elites=False, color=None, hidden=None, hidden_out=None,
             call_dev=None):
        """TDMPC.plan (tdmpc.py:94-163) for a batch of independent problems.

        obs [B,D] (or z0 [B,L]) on the device.  prev_mean [B,H,A] is read (warm start) and
        overwritten with the new mean.  Returns dict(action [B,A], metrics [B,2], prev_mean, taps).
        Everything is enqueued on the current stream; nothing synchronises.
        """
        src = obs if z0 is None else z0
        B = src.shape[0]
        H, I, K, A = int(horizon), int(iterations), int(num_elites), self.A
        ns = list(n_samples) if n_samples is not None else [int(num_samples)] * I
        nr = list(n_reuse) if n_reuse is not None else [0] * I
        npi = list(n_pi) if n_pi is not None else [int(num_pi_trajs)] * I
        Tmax = max(a + b + c for a, b, c in zip(ns, nr, npi))
        Nmax = max(ns)
        self.reserve(B, max(Tmax, int(num_pi_trajs)), H, K)
        if prev_mean is None:
            prev_mean = torch.zeros(B, H, A, device=self.device)
            has_prev = False if has_prev is None else has_prev
        elif has_prev is None:
            has_prev = True
        a = _lib.PlanArgs()
        a.batch, a.horizon, a.iterations, a.num_elites, a.num_pi_trajs = B, H, I, K, int(num_pi_trajs)
        for i in range(I):
            a.n_samples[i], a.n_reuse[i], a.n_pi[i] = ns[i], nr[i], npi[i]
        a.discount, a.momentum = float(discount), float(momentum)
        a.temperature, a.min_std, a.std_floor, a.std_max, a.init_std = temperature, min_std, std_floor, std_max, init_std
        a.eval_mode = int(bool(eval_mode))
        keep = []
        if isinstance(t0, torch.Tensor):
            t0t = t0.to(device=self.device, dtype=torch.uint8).contiguous(); keep.append(t0t)
            a.t0_all, a.t0 = -1, t0t.data_ptr()
        else:
            a.t0_all = int(bool(t0))
        a.has_prev = int(bool(has_prev))
        a.warm_last, a.best_a_last, a.shift_elites = int(bool(warm_last)), int(bool(best_a_last)), int(bool(shift_elites))
        a.elite_horizon = int(elite_horizon)
        if elite_actions is not None:       # [B, K, H, A] in/out (iCEM agent._elite_actions)
            assert elite_actions.is_contiguous() and elite_actions.dtype == torch.float32 and elite_actions.numel() >= B * K * H * A
            a.elite_actions = elite_actions.data_ptr()
        if color is not None:               # dict(len=, mats=[3,len,len] device fp32, n_white=[I], n_color1=[I])
            a.color_len = int(color["len"]); keep.append(color["mats"]); a.color_mats = color["mats"].data_ptr()
            for i in range(I):
                a.n_white[i], a.n_color1[i] = int(color["n_white"][i]), int(color["n_color1"][i])
        if z0 is None:
            obs = _f32(obs, self.device); keep.append(obs); a.obs = obs.data_ptr()
        else:
            z0 = _f32(z0, self.device); keep.append(z0); a.z0 = z0.data_ptr()
        if self.drnn:                       # GRU state in / state after stepping with the returned action out
            if hidden is not None:
                hidden = _f32(hidden, self.device); keep.append(hidden); a.hidden = hidden.data_ptr()
                assert tuple(hidden.shape) == (B, self.Hd)
            if hidden_out is not None:
                assert hidden_out.is_contiguous() and hidden_out.dtype == torch.float32 and tuple(hidden_out.shape) == (B, self.Hd)
                a.hidden_out = hidden_out.data_ptr()
        assert prev_mean.is_contiguous() and prev_mean.dtype == torch.float32 and tuple(prev_mean.shape) == (B, H, A)
        a.prev_mean = prev_mean.data_ptr()
        a.seed = self.seed if seed is None else seed
        if call is None:
            self.call += 1
            call = self.call
        a.call = call
        if call_dev is not None:            # device int64 [1]: added to `call`, incremented at the end of the plan (graph replays)
            assert call_dev.dtype == torch.int64 and call_dev.numel() == 1 and call_dev.is_cuda
            keep.append(call_dev); a.call_dev = call_dev.data_ptr()
        if noise is not None:
            exp = {"pi": (B, H, int(num_pi_trajs), A), "cem": (B, I, H, Nmax, A), "term": (B, I, Tmax, A), "out": (B, A),
                   "shift": (B, H, nr[0], A)}
            for k, fld in (("pi", "noise_pi"), ("cem", "noise_cem"), ("term", "noise_term"), ("out", "noise_out"),
                           ("shift", "noise_shift")):
                if noise.get(k) is not None:
                    t = _f32(noise[k], self.device)
                    assert tuple(t.shape) == exp[k], (k, tuple(t.shape), exp[k])
                    keep.append(t); setattr(a, fld, t.data_ptr())
            if noise.get("u") is not None:
                u = torch.as_tensor(noise["u"], dtype=torch.float64).reshape(B).to(self.device); keep.append(u)
                a.u_pick = u.data_ptr()
        if out is None:
            out = {"action": torch.empty(B, A, device=self.device), "metrics": torch.empty(B, 2, device=self.device)}
        a.action, a.metrics = out["action"].data_ptr(), out["metrics"].data_ptr()
        tp = None
        if taps:
            tp = {"values": torch.zeros(I, B, Tmax, device=self.device),
                  "elite_idx": torch.zeros(I, B, K, dtype=torch.int32, device=self.device),
                  "mean": torch.zeros(I, B, H, A, device=self.device), "std": torch.zeros(I, B, H, A, device=self.device),
                  "z0": torch.zeros(B, self.L, device=self.device)}
            a.tap_values, a.tap_elite_idx = tp["values"].data_ptr(), tp["elite_idx"].data_ptr()
            a.tap_mean, a.tap_std, a.tap_z0 = tp["mean"].data_ptr(), tp["std"].data_ptr(), tp["z0"].data_ptr()
        self._check(self.lib.tdmpc_plan(self._h, C.byref(a), self._stream()), "tdmpc_plan")
        self._keep_call = keep          # inputs must outlive the enqueued kernels
        return {"action": out["action"], "metrics": out["metrics"], "prev_mean": prev_mean, "taps": tp, "hidden": hidden_out}


# ======================================================================================= drop-in
def _fast_plan(self, obs, eval_mode=False, step=None, t0=True):
    """Drop-in for TDMPC.plan (src/algorithm/tdmpc.py:94-163): same arguments, same return
    `(action Tensor[A] on self.device, {'external_reward_mean': float, 'current_std': float})`,
    same `self._prev_mean` side effect, same seed-step early return."""
    cfg = self.cfg
    plan_metrics = {"external_reward_mean": 0.0, "current_std": 0.0}
    if step < cfg.seed_steps and not eval_mode:
        return torch.empty(cfg.action_dim, dtype=torch.float32, device=self.device).uniform_(-1, 1), plan_metrics
    pl = self._planner
    pl.sync_weights(self.model)
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    num_pi_trajs = int(cfg.mixture_coef * cfg.num_samples)
    obs_t = torch.as_tensor(np.asarray(obs), dtype=torch.float32).to(self.device, non_blocking=True).reshape(1, -1)
    prev = getattr(self, "_prev_mean", None)
    # the reference only reads _prev_mean when `not t0` (tdmpc.py:124): at an episode start the horizon may have grown
    # (horizon_schedule) and the stale, shorter history must be ignored, not shape-checked
    has_prev = prev is not None and not t0
    if has_prev and prev.shape[0] != horizon:
        # the reference relies on broadcasting `mean[:-1] = prev[1:]`; only an equal-length (or
        # length-2, i.e. broadcastable) history is meaningful -- emulate it, else fail like torch would
        tail = prev[1:]
        if tail.shape[0] not in (1, horizon - 1):
            raise RuntimeError(f"shape mismatch: _prev_mean has horizon {prev.shape[0]}, plan horizon is {horizon}")
        prev = torch.cat([prev[:1], tail.expand(horizon - 1, -1)], 0)
    buf = prev.reshape(1, horizon, cfg.action_dim).contiguous().clone() if has_prev else None
    noise = getattr(self, "_inject_noise", None)
    res = pl.plan(obs_t, horizon=horizon, num_samples=cfg.num_samples, num_pi_trajs=num_pi_trajs,
                  num_elites=cfg.num_elites, iterations=cfg.iterations, discount=cfg.discount,
                  temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std, std_floor=self.std,
                  eval_mode=eval_mode, t0=bool(t0), prev_mean=buf, has_prev=has_prev, noise=noise,
                  taps=getattr(self, "_want_taps", False))
    self._prev_mean = res["prev_mean"][0]
    self._last_taps = res["taps"]
    m = res["metrics"][0].tolist()          # one D2H read; the reference syncs here too (.item())
    plan_metrics.update({"current_std": m[1], "external_reward_mean": m[0]})
    return res["action"][0], plan_metrics


def icem_sizes(cfg, mixture_coef, have_elites):
    """Per-iteration (sampled, reused, pi) row counts of TdICemSimMlp.plan
    (tdmpc_icem_similarity_mlp.py:182-210): sample-count decay, elite carry-over."""
    out, n = [], cfg.num_samples
    for i in range(cfg.iterations):
        if i > 0:
            n = max(2 * cfg.num_elites, int(n / cfg.factor_decrease_num))
        p = int(mixture_coef * n)
        if p <= 0:
            raise AssertionError("num_pi_trajs must be > 0")            # the reference asserts the same (:184,:203)
        reuse = cfg.fraction_elites_reused > 0 and (have_elites or i > 0)
        out.append((n, int(cfg.fraction_elites_reused * cfg.num_elites) if reuse else 0, p))
    return out


def _fast_plan_icem(self, obs, eval_mode=False, step=None, t0=True):
    """Drop-in for TdICemSimMlp.plan (src/algorithm/tdmpc_icem_similarity_mlp.py:159-265): colored-noise
    sampling, sample-count decay, elite carry-over across steps (shifted) and iterations, best-a, warm
    start with repeated last mean.  Same arguments / return / side effects (plan_horizon, mixture_coef,
    _prev_mean, _elite_actions)."""
    cfg = self.cfg
    plan_metrics = {"external_reward_mean": 0.0, "current_std": 0.0}
    if step < cfg.seed_steps and not eval_mode:
        return torch.empty(cfg.action_dim, dtype=torch.float32, device=self.device).uniform_(-1, 1), plan_metrics
    pl = self._planner
    pl.sync_weights(self.model)
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    extend = False
    if horizon != self.plan_horizon and t0:
        self.plan_horizon = horizon
        extend = True
    H, A, K, I = self.plan_horizon, cfg.action_dim, cfg.num_elites, cfg.iterations
    self.mixture_coef = linear_schedule(cfg.regularization_schedule, step)
    prev_elites = getattr(self, "_elite_actions", None)            # reference layout [H_e, K, A]
    have = prev_elites is not None
    sizes = icem_sizes(cfg, self.mixture_coef, have)
    shift = bool(cfg.shift_elites_over_time) and have
    if have and not shift and sizes[0][1] > 0:
        raise RuntimeError("fraction_elites_reused > 0 without shift_elites_over_time leaves `reused_actions` "
                           "undefined in the reference (tdmpc_icem_similarity_mlp.py:217-231)")
    if cfg.fraction_elites_reused > 0 and not cfg.keep_previous_elites:
        # the reference then keeps feeding iteration 0's `reused_actions` to the later iterations (:224-231); this path
        # always reuses the previous iteration's elites, so refuse rather than diverge silently (as the DRNN path does)
        raise NotImplementedError("fraction_elites_reused > 0 needs keep_previous_elites")
    ebuf = torch.zeros(1, K, H, A, device=self.device)
    eh = H
    if have:
        eh = prev_elites.shape[0]
        if eh not in (H, H - 1) or (eh == H - 1 and not extend):
            raise RuntimeError(f"shape mismatch: _elite_actions has horizon {eh}, plan horizon is {H}")
        ebuf.view(-1)[: K * eh * A] = prev_elites.permute(1, 0, 2).reshape(-1)      # [K][eh][A] at the buffer start
    prev = getattr(self, "_prev_mean", None)
    has_prev = prev is not None and not t0
    if has_prev and prev.shape[0] != H:
        raise RuntimeError(f"shape mismatch: _prev_mean has horizon {prev.shape[0]}, plan horizon is {H}")
    buf = prev.reshape(1, H, A).contiguous().clone() if has_prev else None
    obs_t = torch.as_tensor(np.asarray(obs), dtype=torch.float32).to(self.device, non_blocking=True).reshape(1, -1)
    splits = [mix_split(n) for n, _, _ in sizes]
    key = (cfg.horizon, H, float(cfg.noise_beta))
    if getattr(self, "_color_key", None) != key:
        self._color_mats = torch.from_numpy(color_matrices(cfg.horizon, H, cfg.noise_beta)).to(self.device)
        self._color_key = key
    color = {"len": cfg.horizon, "mats": self._color_mats, "n_white": [s[0] for s in splits], "n_color1": [s[1] for s in splits]}
    res = pl.plan(obs_t, horizon=H, num_samples=cfg.num_samples, num_pi_trajs=sizes[0][2], num_elites=K, iterations=I,
                  discount=cfg.discount, temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std,
                  std_floor=self.std, init_std=0.5, eval_mode=eval_mode, t0=bool(t0), prev_mean=buf, has_prev=has_prev,
                  noise=getattr(self, "_inject_noise", None), taps=getattr(self, "_want_taps", False),
                  n_samples=[s[0] for s in sizes], n_reuse=[s[1] for s in sizes], n_pi=[s[2] for s in sizes],
                  elite_actions=ebuf, elite_horizon=eh, warm_last=True, best_a_last=True, shift_elites=shift, color=color)
    self._prev_mean = res["prev_mean"][0]
    self._elite_actions = ebuf[0].permute(1, 0, 2)                 # [H, K, A] like the reference
    self._last_taps = res["taps"]
    m = res["metrics"][0].tolist()
    plan_metrics.update({"external_reward_mean": m[0], "current_std": m[1]})
    return res["action"][0], plan_metrics


def _fast_plan_icem_drnn(self, obs, hidden, eval_mode=False, step=None, t0=True):
    """Drop-in for TdICemSimDssm.plan (src/algorithm/tdmpc_icem_similarity_drnn.py:168-272): the iCEM loop over the
    GRU latent dynamics.  Same arguments and return `(action [A], hidden' [1, Hd], metrics)`; `hidden'` is the GRU
    state after stepping the model with the returned action (:267).  Differences from the MLP agent that are
    reproduced: single-beta colored noise of length plan_horizon (sample_action_sequence, :135-146) and elite
    reuse gated on `not t0` in every iteration (:203-214)."""
    cfg = self.cfg
    plan_metrics = {"external_reward_mean": 0.0, "current_std": 0.0}
    if step < cfg.seed_steps and not eval_mode:
        return torch.empty(cfg.action_dim, dtype=torch.float32, device=self.device).uniform_(-1, 1), None, plan_metrics
    if cfg.fraction_elites_reused > 0 and not cfg.keep_previous_elites:
        raise NotImplementedError("fraction_elites_reused > 0 needs keep_previous_elites (the reference reuses a stale tensor otherwise)")
    pl = self._planner
    pl.sync_weights(self.model)
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    extend = False
    if horizon != self.plan_horizon and t0:
        self.plan_horizon = horizon
        extend = True
    H, A, K, I = self.plan_horizon, cfg.action_dim, cfg.num_elites, cfg.iterations
    self.mixture_coef = linear_schedule(cfg.regularization_schedule, step)
    prev_elites = getattr(self, "_elite_actions", None)            # reference layout [H_e, K, A]
    have = prev_elites is not None and not t0
    sizes = []
    n = cfg.num_samples
    for i in range(I):
        if i > 0:
            n = max(2 * K, int(n / cfg.factor_decrease_num))
        p = int(self.mixture_coef * n)
        if p <= 0:
            raise AssertionError("num_pi_trajs must be > 0")
        # `hasattr(self, '_elite_actions')` turns true after iteration 0 of this very call; `not t0` gates every iteration
        reuse = cfg.fraction_elites_reused > 0 and not t0 and (prev_elites is not None or i > 0)
        sizes.append((n, int(cfg.fraction_elites_reused * K) if reuse else 0, p))
    shift = bool(cfg.shift_elites_over_time) and have
    if have and not shift and sizes[0][1] > 0:
        raise RuntimeError("fraction_elites_reused > 0 without shift_elites_over_time leaves `reused_actions` undefined "
                           "in the reference (tdmpc_icem_similarity_drnn.py:221-232)")
    ebuf = torch.zeros(1, K, H, A, device=self.device)
    eh = H
    if have:
        eh = prev_elites.shape[0]
        if eh != H:                     # extend_horizon needs t0, and t0 disables reuse: lengths must agree here
            raise RuntimeError(f"shape mismatch: _elite_actions has horizon {eh}, plan horizon is {H}")
        ebuf.view(-1)[: K * eh * A] = prev_elites.permute(1, 0, 2).reshape(-1)
    prev = getattr(self, "_prev_mean", None)
    has_prev = prev is not None and not t0
    if has_prev and prev.shape[0] != H:
        raise RuntimeError(f"shape mismatch: _prev_mean has horizon {prev.shape[0]}, plan horizon is {H}")
    buf = prev.reshape(1, H, A).contiguous().clone() if has_prev else None
    obs_t = torch.as_tensor(np.asarray(obs), dtype=torch.float32).to(self.device, non_blocking=True).reshape(1, -1)
    color = None
    if cfg.noise_beta > 0:
        key = (H, float(cfg.noise_beta))
        if getattr(self, "_color_key", None) != key:
            m = np.zeros((3, H, H), dtype=np.float32)
            m[0] = m[1] = m[2] = synthesis_matrix(cfg.noise_beta, H)
            self._color_mats = torch.from_numpy(m).to(self.device)
            self._color_key = key
        color = {"len": H, "mats": self._color_mats, "n_white": [0] * I, "n_color1": [s[0] for s in sizes]}
    hid = torch.as_tensor(hidden, dtype=torch.float32, device=self.device).reshape(1, -1)
    hid_out = torch.empty(1, cfg.hidden_dim, device=self.device)
    res = pl.plan(obs_t, horizon=H, num_samples=cfg.num_samples, num_pi_trajs=sizes[0][2], num_elites=K, iterations=I,
                  discount=cfg.discount, temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std,
                  std_floor=self.std, init_std=0.5, eval_mode=eval_mode, t0=bool(t0), prev_mean=buf, has_prev=has_prev,
                  noise=getattr(self, "_inject_noise", None), taps=getattr(self, "_want_taps", False),
                  n_samples=[s[0] for s in sizes], n_reuse=[s[1] for s in sizes], n_pi=[s[2] for s in sizes],
                  elite_actions=ebuf, elite_horizon=eh, warm_last=True, best_a_last=True, shift_elites=shift, color=color,
                  hidden=hid, hidden_out=hid_out)
    self._prev_mean = res["prev_mean"][0]
    self._elite_actions = ebuf[0].permute(1, 0, 2)                 # [H, K, A] like the reference
    self._last_taps = res["taps"]
    m = res["metrics"][0].tolist()
    plan_metrics.update({"external_reward_mean": m[0], "current_std": m[1]})
    return res["action"][0], hid_out, plan_metrics


def _fast_plan_sim(self, obs, eval_mode=False, step=None, t0=True, fine_tuning=False):
    """Drop-in for TDMPCSIM.plan (src/algorithm/tdmpc_similarity.py:135-200) -- the agent the quadrotor scripts build
    (train_dyna_episode.py:86,144, train_dyna_episode_multi_agent.py:19).  Differences from TDMPC.plan that are reproduced:
    returns the bare action tensor (no metrics), `fine_tuning` disables the seed-step early return, the planning horizon
    `self.plan_horizon` only moves at an episode start (:151-152), the number of policy trajectories follows
    `regularization_schedule` (:153-154) and may be zero."""
    cfg = self.cfg
    if step < cfg.seed_steps and not eval_mode and not fine_tuning:
        return torch.empty(cfg.action_dim, dtype=torch.float32, device=self.device).uniform_(-1, 1)
    pl = self._planner
    pl.sync_weights(self.model)
    horizon = int(min(cfg.horizon, linear_schedule(cfg.horizon_schedule, step)))
    if horizon != self.plan_horizon and t0:
        self.plan_horizon = horizon
    H = self.plan_horizon
    self.mixture_coef = linear_schedule(cfg.regularization_schedule, step)
    num_pi_trajs = int(self.mixture_coef * cfg.num_samples)
    prev = getattr(self, "_prev_mean", None)
    has_prev = prev is not None and not t0
    if has_prev and prev.shape[0] != H:       # `mean[:-1] = self._prev_mean[1:]` (:165) would fail the same way
        raise RuntimeError(f"shape mismatch: _prev_mean has horizon {prev.shape[0]}, plan horizon is {H}")
    buf = prev.reshape(1, H, cfg.action_dim).contiguous().clone() if has_prev else None
    obs_t = torch.as_tensor(np.asarray(obs), dtype=torch.float32).to(self.device, non_blocking=True).reshape(1, -1)
    res = pl.plan(obs_t, horizon=H, num_samples=cfg.num_samples, num_pi_trajs=num_pi_trajs, num_elites=cfg.num_elites,
                  iterations=cfg.iterations, discount=cfg.discount, temperature=cfg.temperature, momentum=cfg.momentum,
                  min_std=cfg.min_std, std_floor=self.std, eval_mode=eval_mode, t0=bool(t0), prev_mean=buf, has_prev=has_prev,
                  noise=getattr(self, "_inject_noise", None), taps=getattr(self, "_want_taps", False))
    self._prev_mean = res["prev_mean"][0]
    self._last_taps = res["taps"]
    return res["action"][0]


_PLAN_KINDS = {"tdmpc": _fast_plan, "sim": _fast_plan_sim, "icem": _fast_plan_icem, "drnn": _fast_plan_icem_drnn}


def agent_kind(agent):
    """Which of the reference's four planner shapes an agent object has:
    'drnn'  TdICemSimDssm  (GRU dynamics; plan(obs, hidden, ...) -> (a, hidden', metrics))   tdmpc_icem_similarity_drnn.py:81
    'icem'  TdICemSimMlp   (sample_mix_action_sequence; plan -> (a, metrics))                tdmpc_icem_similarity_mlp.py:77
    'sim'   TDMPCSIM       (plan_horizon + fine_tuning; plan -> a)                           tdmpc_similarity.py:74
    'tdmpc' TDMPC          (plan -> (a, metrics))                                            tdmpc.py:53"""
    if hasattr(agent.model._dynamics, "gru_cell"):
        return "drnn"
    if hasattr(type(agent), "sample_mix_action_sequence") or type(agent).__name__ == "TdICemSimMlp":
        return "icem"
    if hasattr(agent, "plan_horizon"):
        return "sim"
    return "tdmpc"


def accelerate(agent, precision="bf16", kind=None, update=False):
    """Give an existing reference-style agent (attributes cfg, model, std, device) the CUDA `plan`.
    The agent's nn.Module keeps owning the parameters; they are re-packed whenever they change.
    `kind` overrides the class detection of agent_kind().  `update=True` (plain TDMPC agents only) also rebinds
    `update(replay_buffer, step)` to the device training step (tdmpc_update): the agent's parameters are stepped in place with
    this library's own Adam state -- agent.optim / agent.pi_optim are no longer used and their state is not carried over."""
    cfg = agent.cfg
    kind = agent_kind(agent) if kind is None else kind
    if kind not in _PLAN_KINDS:
        raise ValueError(f"accelerate: unknown agent kind {kind!r}")
    drnn = kind == "drnn"
    enc = agent.model._encoder
    # helper.enc: Linear, ELU, Linear; helper.dmlab_enc_norm: Linear, <norm>, ELU, Linear with the norm chosen by
    # cfg.norm_type (helper.py:161-165) -- only LayerNorm is implemented; anything else must not be packed as one
    enc_ln = len(enc) == 4
    if enc_ln and not isinstance(enc[1], torch.nn.LayerNorm):
        raise NotImplementedError(f"accelerate: encoder normalisation {type(enc[1]).__name__} is not supported (LayerNorm only)")
    agent._planner = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                             enc_layernorm=enc_ln, precision=precision, device=agent.device,
                             dynamics="drnn" if drnn else "mlp", hidden_dim=cfg.hidden_dim if drnn else 0)
    agent.plan = types.MethodType(_PLAN_KINDS[kind], agent)
    if update:
        if kind != "tdmpc" or enc_ln:
            raise NotImplementedError("accelerate(update=True): the device training step covers the plain TDMPC agent (MLP TOLD, no LayerNorm encoder)")
        for name in ("_train_state", "update_batch", "update"):
            setattr(agent, name, types.MethodType(getattr(TDMPC, name), agent))
    return agent


class TDMPC:
    """Planning half of the reference's TDMPC agent (tdmpc.py:53-163) on the B200 planner."""

    def __init__(self, cfg):
        self.cfg = cfg
        dev = getattr(cfg, "device", "cuda")
        self.device = torch.device(dev if dev != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.std = linear_schedule(cfg.std_schedule, 0)
        self.model = TOLD(cfg).to(self.device)
        self.model_target = deepcopy(self.model)
        self.model.eval()
        self.model_target.eval()
        self._planner = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                                enc_layernorm=getattr(cfg, "enc_layernorm", False),
                                precision=getattr(cfg, "precision", "bf16"), device=self.device)

    plan = _fast_plan

    def state_dict(self):
        return {"model": self.model.state_dict(), "model_target": self.model_target.state_dict()}

    def save(self, fp):
        torch.save(self.state_dict(), fp)

    def load(self, fp):
        d = torch.load(fp)
        self.model.load_state_dict(d["model"])
        self.model_target.load_state_dict(d["model_target"])

    @torch.no_grad()
    def estimate_value(self, z, actions, horizon):
        """Reference signature (tdmpc.py:83-92): z [T,L], actions [H,T,A] -> (G [T,1], float)."""
        self._planner.sync_weights(self.model)
        # one "problem" per trajectory so every row may start from its own latent
        acts = actions[:horizon].permute(1, 0, 2).unsqueeze(1).contiguous()      # [T, 1, H, A]
        pl = self._planner
        pl.call += 1                    # fresh terminal-policy noise per call, like the reference's _standard_normal draw
        v, r = pl.estimate_value(z, acts, discount=self.cfg.discount, min_std=self.cfg.min_std, seed=pl.seed, call=pl.call)
        return v, float(r.mean().item())

    # ------------------------------------------------------------------------------------------------ training step
    TRAIN_KEYS = tuple(f"{m}.{i}.{s}" for m, idx in (("_encoder", (0, 2)), ("_dynamics", (0, 2, 4)), ("_reward", (0, 2, 4)), ("_pi", (0, 2, 4)))
                       for i in idx for s in ("weight", "bias")) + \
        tuple(f"{q}.{i}.{s}" for q in ("_Q1", "_Q2") for i in (0, 1, 3, 4, 6) for s in ("weight", "bias"))

    @staticmethod
    def _train_weights(model, device):
        """tdmpc_weights over the module's OWN parameter storage (the training step updates it in place)."""
        sd = dict(model.state_dict(keep_vars=True))
        w = _lib.Weights()
        for k in TDMPC.TRAIN_KEYS:
            t = sd[k]
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != device:
                raise ValueError(f"update(): parameter {k} must be a contiguous fp32 tensor on {device}")

        def P(k):
            return sd[k].data_ptr()

        w.enc_w[0], w.enc_b[0], w.enc_w[1], w.enc_b[1] = P("_encoder.0.weight"), P("_encoder.0.bias"), P("_encoder.2.weight"), P("_encoder.2.bias")
        for dst_w, dst_b, mod in ((w.dyn_w, w.dyn_b, "_dynamics"), (w.rew_w, w.rew_b, "_reward"), (w.pi_w, w.pi_b, "_pi")):
            for j, i in enumerate((0, 2, 4)):
                dst_w[j], dst_b[j] = P(f"{mod}.{i}.weight"), P(f"{mod}.{i}.bias")
        for h, q in enumerate(("_Q1", "_Q2")):
            for j, i in enumerate((0, 3, 6)):
                w.q_w[h][j], w.q_b[h][j] = P(f"{q}.{i}.weight"), P(f"{q}.{i}.bias")
            for j, i in enumerate((1, 4)):
                w.q_ln_g[h][j], w.q_ln_b[h][j] = P(f"{q}.{i}.weight"), P(f"{q}.{i}.bias")
        return w

    def _train_state(self, B, H):
        """Flat optimiser arenas (state_dict order; tdmpc_train_args) and the workspace, created on first use."""
        pl = self._planner
        st = getattr(self, "_train", None)
        if st is None:
            lo, hi = C.c_int64(), C.c_int64()
            n = pl.lib.tdmpc_train_param_count(C.byref(pl._cfg), C.byref(lo), C.byref(hi))
            if n != sum(p.numel() for p in self.model.parameters()):
                raise RuntimeError("update(): the model is not the plain MLP TOLD the training step supports")
            z = lambda k: torch.zeros(k, device=self.device)
            st = self._train = {"n": n, "pi": (lo.value, hi.value), "m": z(n), "v": z(n), "pi_m": z(hi.value - lo.value), "pi_v": z(hi.value - lo.value),
                                "grads": z(n), "step": 0, "pi_step": 0, "metrics": z(8), "ws": None, "ws_key": None}
        if st["ws_key"] != (B, H):
            nbytes = pl.lib.tdmpc_train_workspace_bytes(C.byref(pl._cfg), B, H)
            st["ws"] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            st["ws_key"] = (B, H)
        return st

    @torch.no_grad()
    def update_batch(self, obs, next_obses, action, reward, weights, step, noise_pi=None, noise_pi_update=None):
        """One whole training step (tdmpc.py:191-245) on a replay batch, on the device (tdmpc_update): forward, backward, gradient
        clipping, Adam over the model, update_pi with its own Adam, and the target EMA when `step % update_freq == 0`.
        self.model / self.model_target are updated in place.  Returns the reference's metric dict plus `priorities` [B,1]
        (device) and `pi_grad_norm`."""
        cfg, pl, dev = self.cfg, self._planner, self.device
        H = int(cfg.horizon)
        obs = _f32(obs, dev); B = obs.shape[0]
        nxt, act = _f32(next_obses[:H], dev), _f32(action[:H], dev)
        rew = _f32(reward[:H], dev).reshape(H, B)
        wts = _f32(weights, dev).reshape(B)
        st = self._train_state(B, H)
        pri = torch.empty(B, device=dev)
        a = _lib.TrainArgs()
        a.online, a.target = TDMPC._train_weights(self.model, dev), TDMPC._train_weights(self.model_target, dev)
        a.batch, a.horizon, a.discount, a.rho, a.min_std = B, H, float(cfg.discount), float(cfg.rho), float(cfg.min_std)
        a.consistency_coef, a.reward_coef, a.value_coef = float(cfg.consistency_coef), float(cfg.reward_coef), float(cfg.value_coef)
        a.lr, a.beta1, a.beta2, a.adam_eps = float(cfg.lr), 0.9, 0.999, 1e-8
        a.grad_clip_norm, a.tau = float(cfg.grad_clip_norm), float(cfg.tau)
        a.update_target = 1 if step % cfg.update_freq == 0 else 0
        a.adam_step, a.pi_adam_step = st["step"], st["pi_step"]
        a.adam_m, a.adam_v, a.pi_adam_m, a.pi_adam_v, a.grads = (st[k].data_ptr() for k in ("m", "v", "pi_m", "pi_v", "grads"))
        a.obs, a.next_obses, a.action, a.reward, a.weights = obs.data_ptr(), nxt.data_ptr(), act.data_ptr(), rew.data_ptr(), wts.data_ptr()
        keep = [obs, nxt, act, rew, wts]
        if noise_pi is not None:
            n1 = _f32(noise_pi, dev); assert tuple(n1.shape) == (H, B, cfg.action_dim)
            a.noise_pi = n1.data_ptr(); keep.append(n1)
        if noise_pi_update is not None:
            n2 = _f32(noise_pi_update, dev); assert tuple(n2.shape) == (H + 1, B, cfg.action_dim)
            a.noise_pi_update = n2.data_ptr(); keep.append(n2)
        pl.call += 1
        a.seed, a.call = pl.seed, pl.call
        a.metrics, a.priorities = st["metrics"].data_ptr(), pri.data_ptr()
        a.workspace, a.workspace_bytes = st["ws"].data_ptr(), st["ws"].numel()
        pl._check(pl.lib.tdmpc_update(pl._h, C.byref(a), pl._stream()), "tdmpc_update")
        st["step"] += 1; st["pi_step"] += 1
        pl._wver = None                                   # the parameters moved: re-pack before the next plan()
        m = st["metrics"].tolist()                        # the one device -> host read of the step
        self._keep_train = keep
        return {"consistency_loss": m[0], "reward_loss": m[1], "value_loss": m[2], "pi_loss": m[6], "total_loss": m[3],
                "weighted_loss": m[4], "grad_norm": m[5], "pi_grad_norm": m[7], "priorities": pri.unsqueeze(1)}

    def update(self, replay_buffer, step):
        """TDMPC.update (tdmpc.py:191-245): same call, same returned metric names; every tensor operation of it runs in
        tdmpc_update (csrc/train.cu) on the module's own parameters."""
        obs, next_obses, action, reward, idxs, weights = replay_buffer.sample()
        self.std = linear_schedule(self.cfg.std_schedule, step)
        m = self.update_batch(obs, next_obses, action, reward, weights, step)
        replay_buffer.update_priorities(idxs, m.pop("priorities"))
        m.pop("pi_grad_norm")
        return m

    @torch.no_grad()
    def update_forward(self, obs, next_obses, action, reward, weights=None, noise_pi=None, taps=False):
        """Forward half of TDMPC.update (tdmpc.py:191-228) for one replay batch, as CUDA (tdmpc_update_forward, fp32):
        obs [B,D], next_obses [>=H,B,D], action [>=H,B,A], reward [>=H,B,1] or [>=H,B] -- what ReplayBuffer.sample() returns --
        and `weights` [B,1] its importance weights.  Returns the reference's metric names (consistency_loss, reward_loss,
        value_loss, total_loss, weighted_loss: batch means as floats), the per-sample `priorities` [B,1] that update() hands to
        update_priorities (already clamped at 1e4), the per-sample loss tensors and the latents `zs` [H+1,B,L] that update_pi
        consumes.  Backward and the optimiser step are not part of this call."""
        cfg = self.cfg
        pl = self._planner
        pl.sync_weights(self.model)
        if getattr(self, "_planner_target", None) is None:
            self._planner_target = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                                           enc_layernorm="_encoder.3.weight" in self.model_target.state_dict(),
                                           precision="fp32", device=self.device)
        pt = self._planner_target
        # always re-packed: the reference moves the target with `p_target.data.lerp_(p.data, tau)` (helper.ema, helper.py:48-52),
        # and writes through `.data` do not bump Tensor._version, which is what sync_weights watches
        pt.sync_weights(self.model_target, force=True)
        H = int(cfg.horizon)
        dev = self.device
        obs = _f32(obs, dev); B = obs.shape[0]
        nxt, act = _f32(next_obses[:H], dev), _f32(action[:H], dev)
        rew = _f32(reward[:H], dev).reshape(H, B)
        out = {k: torch.empty(B, device=dev) for k in ("consistency", "reward", "value", "priority", "total")}
        zs = torch.empty(H + 1, B, cfg.latent_dim, device=dev)
        a = _lib.UpdateArgs()
        a.batch, a.horizon, a.discount, a.rho, a.min_std = B, H, float(cfg.discount), float(cfg.rho), float(cfg.min_std)
        a.consistency_coef, a.reward_coef, a.value_coef = float(cfg.consistency_coef), float(cfg.reward_coef), float(cfg.value_coef)
        a.obs, a.next_obses, a.action, a.reward = obs.data_ptr(), nxt.data_ptr(), act.data_ptr(), rew.data_ptr()
        nz = None
        if noise_pi is not None:
            nz = _f32(noise_pi, dev)
            assert tuple(nz.shape) == (H, B, cfg.action_dim)
            a.noise_pi = nz.data_ptr()
        pl.call += 1
        a.seed, a.call = pl.seed, pl.call
        a.consistency_loss, a.reward_loss, a.value_loss = out["consistency"].data_ptr(), out["reward"].data_ptr(), out["value"].data_ptr()
        a.priority_loss, a.total_loss, a.zs = out["priority"].data_ptr(), out["total"].data_ptr(), zs.data_ptr()
        tp = None
        if taps:
            tp = {k: torch.empty(H, B, device=dev) for k in ("td_target", "q1", "q2", "reward_pred")}
            a.td_target, a.q1, a.q2, a.reward_pred = (tp[k].data_ptr() for k in ("td_target", "q1", "q2", "reward_pred"))
        pl._check(pl.lib.tdmpc_update_forward(pl._h, pt._h, C.byref(a), pl._stream()), "tdmpc_update_forward")
        self._keep_update = (obs, nxt, act, rew, nz)
        w = torch.ones(B, device=dev) if weights is None else _f32(weights, dev).reshape(B)
        return {"consistency_loss": float(out["consistency"].mean().item()), "reward_loss": float(out["reward"].mean().item()),
                "value_loss": float(out["value"].mean().item()), "total_loss": float(out["total"].mean().item()),
                "weighted_loss": float((out["total"] * w).mean().item()),
                "priorities": out["priority"].clamp(max=1e4).unsqueeze(1), "per_sample": out, "zs": zs, "taps": tp}


class TdICemSimMlp(TDMPC):
    """Planning half of the reference's iCEM agent with MLP dynamics
    (src/algorithm/tdmpc_icem_similarity_mlp.py:77-265) on the B200 planner."""

    def __init__(self, cfg):
        self.cfg = cfg
        dev = getattr(cfg, "device", "cuda")
        self.device = torch.device(dev if dev != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.std = linear_schedule(cfg.std_schedule, 0)
        self.mixture_coef = linear_schedule(cfg.regularization_schedule, 0)
        self.model = DSSM(cfg).to(self.device)
        self.model_target = deepcopy(self.model)
        self.model.eval()
        self.model_target.eval()
        self.plan_horizon = 1
        self._planner = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                                enc_layernorm=bool(cfg.normalize), precision=getattr(cfg, "precision", "bf16"),
                                device=self.device)

    plan = _fast_plan_icem


class TdICemSimDssm(TDMPC):
    """Planning half of the reference's iCEM agent with GRU latent dynamics
    (src/algorithm/tdmpc_icem_similarity_drnn.py:80-272) on the B200 planner: `plan(obs, hidden, eval_mode, step, t0)`
    -> `(action, hidden', metrics)`."""

    def __init__(self, cfg):
        self.cfg = cfg
        dev = getattr(cfg, "device", "cuda")
        self.device = torch.device(dev if dev != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.std = linear_schedule(cfg.std_schedule, 0)
        self.mixture_coef = linear_schedule(cfg.regularization_schedule, 0)
        self.model = DSSMRnn(cfg).to(self.device)
        self.model_target = deepcopy(self.model)
        self.model.eval()
        self.model_target.eval()
        self.plan_horizon = 1
        self.explore_coef = 0
        self._planner = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                                enc_layernorm=bool(cfg.normalize), precision=getattr(cfg, "precision", "fp32"),
                                device=self.device, dynamics="drnn", hidden_dim=cfg.hidden_dim)

    plan = _fast_plan_icem_drnn

    @torch.no_grad()
    def estimate_value(self, z, actions, hidden):
        """Reference signature (tdmpc_icem_similarity_drnn.py:124-133): z [T,L], actions [H,T,A], hidden [T,Hd] -> (G [T,1], float)."""
        self._planner.sync_weights(self.model)
        H = self.plan_horizon
        acts = actions[:H].permute(1, 0, 2).unsqueeze(1).contiguous()      # one "problem" per trajectory: [T, 1, H, A]
        pl = self._planner
        pl.call += 1
        v, r = pl.estimate_value(z, acts, discount=self.cfg.discount, min_std=self.cfg.min_std, hidden=hidden, seed=pl.seed, call=pl.call)
        return v, float(r.mean().item())


class TDMPCSIM(TDMPC):
    """Planning half of the reference's TDMPCSIM agent (src/algorithm/tdmpc_similarity.py:74-200) on the B200 planner:
    `plan(obs, eval_mode, step, t0, fine_tuning) -> action`."""

    def __init__(self, cfg):
        self.cfg = cfg
        dev = getattr(cfg, "device", "cuda")
        self.device = torch.device(dev if dev != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.std = linear_schedule(cfg.std_schedule, 0)
        self.mixture_coef = linear_schedule(cfg.regularization_schedule, 0)
        self.model = DSSM(cfg).to(self.device)            # same sub-modules / state_dict keys as tdmpc_similarity.TOLD
        self.model_target = deepcopy(self.model)
        self.model.eval()
        self.model_target.eval()
        self.plan_horizon = 1
        self.batch_size = getattr(cfg, "batch_size", None)
        self._planner = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim,
                                enc_layernorm=bool(cfg.normalize), precision=getattr(cfg, "precision", "bf16"),
                                device=self.device)

    plan = _fast_plan_sim

    @torch.no_grad()
    def estimate_value(self, z, actions, horizon):
        """Reference signature (tdmpc_similarity.py:124-133): returns G [T,1] only."""
        return TDMPC.estimate_value(self, z, actions, horizon)[0]

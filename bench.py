#!/usr/bin/env python
"""Benchmark of the TD-MPC planning hot path (BASELINE.json metric: plans/sec and rollout
sample-steps/sec at 1/2/4/8 B200 next to the host-CPU plan()).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--batch B] [--precision bf16]
  python bench.py --impl reference ...     # the reference's CPU plan() (oracle port) on the host cores

A "step" is one TDMPC.plan() over one batch of independent environments (synthetic observations,
random-init TOLD weights with re-randomised heads).  Default workload = BASELINE.json configs[2] AS WRITTEN:
quadrotor batched planning, 4096 environment instances in total, SHARDED over the N GPUs (4096 / N per rank,
contiguous blocks) -- strong scaling.  One process per GPU (torchrun), no data-path collective, one all-gather
of the planned actions per step.  `--batch B` overrides the TOTAL number of environments.  The weak-scaling line
(4096 environments per GPU) and, at N = 8, BASELINE configs[4] (65,536 environments x 1024 samples through the
GRU dynamics, 8192 per GPU) are measured in the same run and reported under "extra"; configs[0] / [1]
(single-environment cartpole / dog-run) are latency cases, also under "extra", each with its own roofline and
CPU / eager-GPU baselines.

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402  (deterministic synthetic weights / observations)

WORKLOADS = {
    # name: (cfg name in tests/synth.py, default batch per GPU, description)
    "cfg3": ("cfg3", 4096, "quadrotor batched planning: obs 18, act 4, L=50, M=512, H=5, N=512, P=25, K=64, I=6"),
    "cfg1": ("cfg1", 1, "TD-MPC state default (cartpole dims): obs 5, act 1, L=50, M=512, H=5, N=512, P=25, K=64, I=6"),
    "cfg2": ("cfg2", 1, "dog-run dims: obs 223, act 38, L=100, M=512, H=5, N=512, P=25, K=64, I=8"),
    "cfg2s": ("cfg2s", 1, "dog-run dims with L=50, I=6"),
}


def flops_and_steps(cfg):
    from tdmpc_b200.accounting import flops_per_plan, sample_steps_per_plan
    H, N, P, T = synth.plan_dims(cfg)
    rows = [T] * cfg.iterations
    f = flops_per_plan(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, H, rows, P)
    return f, sample_steps_per_plan(H, rows, P), (H, N, P, T)


def workload_config(args, world, B_total, T):
    """The `config` object of the JSON line -- identical for the b200 arm and the reference arm."""
    _, _, desc = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: {desc}", "envs_total": B_total, "envs_per_gpu": B_total // world,
            "trajectories_per_plan": T, "parallelism": f"dp{world} over independent planning problems (contiguous shards of the environments)",
            "l2": "flushed (256 MB write) between timed iterations",
            "weights": "random-init TOLD, heads re-randomised (tests/synth.py)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def icem_batched(dev, precision, B, drnn, reps=3, roofline=False):
    """Batched iCEM planning through Planner.plan (H=10, N=1024 decaying by 1.25, K=32, mixture 0.5, I=6, obs 18, act 4): elite
    carry-over, colored noise, best-a.  drnn=False: TdICemSimMlp (BASELINE.json configs[3]: MLP dynamics, white/pink/brown mix);
    drnn=True: TdICemSimDssm (configs[4]: GRU dynamics Hd=128, single-beta colored noise, hidden state in/out)."""
    from tdmpc_b200 import Planner
    from tdmpc_b200.colored import color_matrices, mix_split, synthesis_matrix
    cfg = synth.make_cfg("cfg5" if drnn else "cfg4")
    extra_kw = dict(dynamics="drnn", hidden_dim=cfg.hidden_dim) if drnn else {}
    pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, enc_layernorm=True, precision=precision,
                 device=dev, **extra_kw)
    sd = synth.make_drnn_weights(cfg, 0) if drnn else synth.make_told_weights(cfg, 0, enc_ln=True)
    pl.sync_weights({k: v.to(dev) for k, v in sd.items()})
    H, K, I, A = cfg.horizon, cfg.num_elites, cfg.iterations, cfg.action_dim
    sizes, n = [], cfg.num_samples
    for i in range(I):
        if i > 0:
            n = max(2 * K, int(n / cfg.factor_decrease_num))
        sizes.append((n, int(cfg.fraction_elites_reused * K), int(0.5 * n)))
    if drnn:
        mats = np.zeros((3, H, H), dtype=np.float32)
        mats[0] = mats[1] = mats[2] = synthesis_matrix(cfg.noise_beta, H)
        color = {"len": H, "mats": torch.from_numpy(mats).to(dev), "n_white": [0] * I, "n_color1": [s[0] for s in sizes]}
    else:
        splits = [mix_split(s[0]) for s in sizes]
        color = {"len": H, "mats": torch.from_numpy(color_matrices(H, H, cfg.noise_beta)).to(dev),
                 "n_white": [sp[0] for sp in splits], "n_color1": [sp[1] for sp in splits]}
    obs = torch.from_numpy(synth.make_obs(cfg, 1, batch=B)).to(dev)
    kw = dict(horizon=H, num_samples=cfg.num_samples, num_pi_trajs=sizes[0][2], num_elites=K, iterations=I, discount=cfg.discount,
              temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std, std_floor=cfg.min_std, init_std=0.5, eval_mode=True,
              n_samples=[s[0] for s in sizes], n_reuse=[s[1] for s in sizes], n_pi=[s[2] for s in sizes],
              elite_actions=torch.zeros(B, K, H, A, device=dev), elite_horizon=H, warm_last=True, best_a_last=True, shift_elites=True,
              color=color, prev_mean=torch.zeros(B, H, A, device=dev))
    if drnn:
        hid = torch.zeros(B, cfg.hidden_dim, device=dev)
        kw.update(hidden=hid, hidden_out=torch.empty_like(hid))
    for _ in range(2):
        pl.plan(obs, t0=False, has_prev=True, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pl.plan(obs, t0=False, has_prev=True, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows = [sum(s) for s in sizes]
    from tdmpc_b200.accounting import flops_per_plan, value_launch_flops
    L, M, Hd = cfg.latent_dim, cfg.mlp_dim, cfg.hidden_dim if drnn else 0
    flop = flops_per_plan(cfg.obs_shape[0], A, L, M, cfg.enc_dim, H, rows, sizes[0][2], Hd)
    out = {"workload": f"iCEM + {'GRU' if drnn else 'MLP'} dynamics, {B} environments per launch, rows per iteration {rows}, precision {precision}",
           "ms_per_batched_plan": ms, "plans_per_sec": B / ms * 1e3, "sample_steps_per_sec": B / ms * 1e3 * (sum(rows) + sizes[0][2]) * H,
           "flop_per_plan": flop, "tflops": B / ms * 1e3 * flop / 1e12}
    if roofline:
        # the dominant kernel alone, at the first iteration's shape (value-mode launch over B * T_0 trajectories)
        sus, burst, _, src = measured_peaks()
        z0 = pl.encode(obs)
        acts = torch.rand(B, rows[0], H, A, device=dev) * 2 - 1
        ekw = dict(discount=cfg.discount, min_std=cfg.min_std, seed=1, call=1)
        if drnn:
            ekw["hidden"] = torch.zeros(B, cfg.hidden_dim, device=dev)
        k_ms = time_kernel(lambda: pl.estimate_value(z0, acts, **ekw))
        fl = value_launch_flops(B, rows[0], H, A, L, M, Hd)
        out["roofline"] = {"bound": "tensor", "kernel": f"rollout_tc_kernel<{'GRU' if drnn else 'MLP'}> (value mode)", "achieved": fl / (k_ms * 1e-3) / 1e12,
                           "peak": sus, "unit": "TFLOP/s", "frac": fl / (k_ms * 1e-3) / 1e12 / sus, "traffic": None, "kernel_ms": k_ms,
                           "flop_per_launch": fl, "peak_source": f"{src}: bf16_tflops_sustained"}
    return out


_FLUSH = {}


def time_kernel(fn, reps=3):
    """Mean CUDA-event time of fn() on the current stream, L2 flushed (256 MB write) before every timed call."""
    dev = torch.cuda.current_device()
    if dev not in _FLUSH:
        _FLUSH[dev] = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{dev}")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt = []
    for _ in range(reps):
        _FLUSH[dev].fill_(1)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        kt.append(e0.elapsed_time(e1))
    return float(np.mean(kt))


def drnn_batched(dev, precision, B, **kw):
    return icem_batched(dev, precision, B, True, **kw)


def replay_extra(dev, hbm_peak):
    """ReplayBuffer.sample at the reference's full size (max_buffer_size 1e6 priorities, batch 512,
    horizon 5, quadrotor dims).  Device time by CUDA events around tdmpc_replay_sample (+ gather), against the reference's
    way of doing it restated on the host (oracle port: fp32 pow / sum, float64 cumsum + searchsorted, gathers)."""
    from oracle import replay_oracle as RO
    from tdmpc_b200 import ReplayBuffer
    cfg = synth.make_replay_cfg("small", train_steps=1000000, episode_length=1000, batch_size=512, obs_shape=(18,), action_dim=4)
    buf = ReplayBuffer(cfg, latent_plan=True)
    rs = np.random.RandomState(7)
    pri = np.exp(1.5 * rs.randn(buf.capacity)).astype(np.float32)
    pri[(np.arange(buf.capacity) % 1000) >= 995] = 0
    buf._priorities.copy_(torch.from_numpy(pri))
    buf._obs.normal_(); buf._action.uniform_(-1, 1); buf._reward.normal_()
    out = {"workload": "ReplayBuffer.sample: 1e6 priorities, batch 512, horizon 5, obs 18, act 4, per_alpha 0.6, per_beta 0.4"}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for full in (False, True):
        buf.idx, buf._full = (0, True) if full else (buf.capacity - 1000, False)
        total = buf.capacity if full else buf.idx
        for _ in range(3):
            buf.sample()
        ms_w = []
        for _ in range(10):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            buf.sample()
            torch.cuda.synchronize()
            ms_w.append((time.perf_counter() - t0) * 1e3)
        # device time of the draw alone: 40 calls back to back over 40 different priority vectors (160 MB > L2, so every call
        # reads HBM), one event pair around the lot; the host enqueues ahead of the GPU, no status read in between
        many = torch.from_numpy(pri).to(dev).repeat(40, 1)
        idxs = torch.empty(512, dtype=torch.int64, device=dev); wts = torch.empty(512, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        calls = [buf._sample_args(idxs, wts, priorities=many[i]) for i in range(40)]
        ms_k = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            e0.record()
            for a in calls:
                assert buf._lib.tdmpc_replay_sample(ctypes.byref(a), stream) == 0
            e1.record(); torch.cuda.synchronize()
            ms_k.append(e0.elapsed_time(e1) / 40)
        assert buf._status.tolist()[0] in (512,) or full
        del many
        rounds = buf.last_rounds
        alg = 2 * 4 * total + rounds * (4 + int(full)) * total      # pass 1 reads the priorities and writes their powers; per round: block sums (+ mask)
        k = float(np.median(ms_k))
        key = "replace_false_full_buffer" if full else "replace_true"
        out[key] = {"ms_sample_indices_device": k, "ms_sample_wall_incl_gather_and_status_read": float(np.median(ms_w)),
                    "rounds": rounds, "launches": 1 + 2 * (-(-rounds // ReplayBuffer.ROUNDS_PER_CALL) * ReplayBuffer.ROUNDS_PER_CALL if full else 1) + 1,
                    "algorithmic_bytes": alg, "achieved_gb_s": alg / (k * 1e-3) / 1e9,
                    "frac_of_hbm_peak": alg / (k * 1e-3) / 1e9 / hbm_peak, "hbm_peak_gb_s": hbm_peak}
    st = RO.ReplayState(cfg, latent_plan=True)
    st["priorities"] = torch.from_numpy(pri); st["idx"] = st["capacity"] - 1000
    st["obs"].normal_()
    t_cpu = []
    for i in range(5):
        uni = np.random.RandomState(i).random_sample(512)
        t0 = time.perf_counter()
        RO.sample(st, uni)
        t_cpu.append((time.perf_counter() - t0) * 1e3)
    out["cpu_port_ms"] = float(np.median(t_cpu))
    out["cpu_port_note"] = (f"oracle/replay_oracle.sample on {torch.get_num_threads()} host threads, buffers already in host memory "
                            "(the reference additionally copies the 4 MB priority vector device -> host per call, helper.py:511)")
    return out


def update_forward_extra(dev, batch=512):
    """SURVEY 8(f) row 2, forward half: TDMPC.update's loss terms for one replay batch (cfg3 dims, batch 512, H = 5) through
    tdmpc_update_forward (fp32 CUDA-core kernel), next to the same op sequence run eagerly by torch on this GPU and on the host."""
    from oracle import tdmpc_oracle as O
    from tdmpc_b200 import TDMPC
    cfg = synth.make_cfg("cfg3", batch_size=batch, device=str(dev), precision="fp32")
    sd, sd_t = synth.make_told_weights(cfg, 0), synth.make_told_weights(cfg, 50)
    agent = TDMPC(cfg)
    agent.model.load_state_dict(sd); agent.model_target.load_state_dict(sd_t)
    b = {k: v.to(dev) for k, v in synth.make_update_batch(cfg, 0, batch).items()}
    args = (b["obs"], b["next_obses"], b["action"], b["reward"])
    ms = time_kernel(lambda: agent.update_forward(*args, weights=b["weights"]), reps=5)
    sdd, sdt = {k: v.to(dev) for k, v in sd.items()}, {k: v.to(dev) for k, v in sd_t.items()}

    def eager():
        with torch.device(dev):
            O.update_forward(sdd, sdt, cfg, b)
    ms_e = time_kernel(eager, reps=5)
    bc = synth.make_update_batch(cfg, 0, batch)
    t = []
    for _ in range(4):
        t0 = time.perf_counter(); O.update_forward(sd, sd_t, cfg, bc); t.append((time.perf_counter() - t0) * 1e3)
    L, A, M = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim
    from tdmpc_b200.accounting import mlp_flops
    f_dyn, f_rew, f_pi, f_q = mlp_flops(L, A, M)
    f_enc = 2 * (cfg.obs_shape[0] * cfg.enc_dim + cfg.enc_dim * L)
    flop = batch * (f_enc + cfg.horizon * (f_dyn + f_rew + 2 * f_q + 2 * f_enc + f_pi + 2 * f_q))
    return {"workload": f"TDMPC.update forward (losses, TD targets, priorities): cfg3 dims, batch {batch}, H = {cfg.horizon}, fp32",
            "ms_per_call": ms, "incl": "weight-change check + target re-pack + batch-mean reductions (host syncs)", "flop_per_call": flop,
            "tflops_fp32": flop / (ms * 1e-3) / 1e12, "gpu_eager_ms": ms_e, "cpu_port_ms": float(np.median(t)), "cpu_threads": torch.get_num_threads()}


def train_step_extra(dev, batch=512):
    """SURVEY 8(f) row 2, whole: one TDMPC.update (forward, backward through time, clip, both Adams, update_pi, ema) through
    tdmpc_update (fp32 in / fp32 out; the GEMMs as split-bf16 tcgen05 MMAs with fp32 accumulation, replayed as one CUDA graph)
    next to the same step run eagerly by torch on this GPU (autograd + the oracle's op sequence) and on
    the host cores."""
    from oracle import tdmpc_oracle as O
    from tdmpc_b200 import TDMPC
    from tdmpc_b200.accounting import mlp_flops
    cfg = synth.make_cfg("cfg3", batch_size=batch, device=str(dev), precision="fp32")
    sd, sd_t = synth.make_told_weights(cfg, 0), synth.make_told_weights(cfg, 50)
    agent = TDMPC(cfg)
    agent.model.load_state_dict(sd); agent.model_target.load_state_dict(sd_t)
    b = {k: v.to(dev) for k, v in synth.make_train_batch(cfg, 0, batch).items()}
    n0 = agent._planner.launch_count
    step = [0]

    def ours():
        agent.update_batch(b["obs"], b["next_obses"], b["action"], b["reward"], b["weights"], step[0]); step[0] += 1
    ms = time_kernel(ours, reps=10)
    launches = (agent._planner.launch_count - n0) / step[0]
    st = {"sd": {k: v.to(dev) for k, v in sd.items()}, "t": {k: v.to(dev) for k, v in sd_t.items()}, "opt": None}

    def eager():
        with torch.device(dev):
            _, st["sd"], st["t"], _, st["opt"] = O.update_step(st["sd"], st["t"], cfg, b, st["opt"], True)
    ms_e = time_kernel(eager, reps=3)
    bc = synth.make_train_batch(cfg, 0, batch)
    t0 = time.perf_counter(); O.update_step(sd, sd_t, cfg, bc, None, True); ms_c = (time.perf_counter() - t0) * 1e3
    L, A, M, H = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim, cfg.horizon
    f_dyn, f_rew, f_pi, f_q = mlp_flops(L, A, M)
    f_enc = 2 * (cfg.obs_shape[0] * cfg.enc_dim + cfg.enc_dim * L)
    fwd = f_enc + H * (f_dyn + f_rew + 2 * f_q)                       # online model, differentiated: backward ~ 2x
    tgt = H * (2 * f_enc + f_pi + 2 * f_q)                            # TD targets (no gradient)
    upi = (H + 1) * (f_pi + 2 * f_q)                                  # update_pi forward; backward: 2x policy, 1x Q (data gradients only)
    flop = batch * (3 * fwd + tgt + upi + (H + 1) * (2 * f_pi + 2 * f_q))
    return {"workload": f"TDMPC.update whole (forward, backward, clip, Adam x2, update_pi, ema): cfg3 dims, batch {batch}, H = {H}, fp32 "
                        "(GEMMs: hi/lo bf16 split, 3 tcgen05 MMAs per product, fp32 TMEM accumulators; one CUDA graph per step)",
            "ms_per_step": ms, "launches_per_step": launches, "flop_per_step": flop, "tflops_fp32": flop / (ms * 1e-3) / 1e12,
            "gpu_eager_ms": ms_e, "cpu_port_ms": ms_c, "cpu_threads": torch.get_num_threads(),
            "roofline": {"bound": "tensor", "kernel": "gemm_tc::gemm_kernel (91 of the step's ~270 launches), whole step timed",
                         "achieved": flop / (ms * 1e-3) / 1e12, "peak": measured_peaks()[0], "unit": "TFLOP/s",
                         "frac": flop / (ms * 1e-3) / 1e12 / measured_peaks()[0], "traffic": None,
                         "note": "useful (fp32-equivalent) flops of the whole step over its wall time; the tensor pipe executes 3x that "
                                 "(hi/lo bf16 split).  The step is latency-bound -- ~270 dependent graph nodes x 7 us, tensor pipe 5-20 % "
                                 "active inside the GEMM (profiles/r2d_gemm_tc_ncu_summary.json) -- not tensor- or HBM-bound"},
            "d2h_bytes_per_step": 32, "note": "parameters of the torch modules are updated in place; one 8-float metrics read per step"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1420.0), d.get("bf16_tflops", 1701.7), d.get("hbm_gbs", 6448.7), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
def _oracle_planner(kind, cfg, sd, device=None):
    """(make_inputs(i) -> inputs, run(inputs, i)) for the oracle port of one of the reference planners.  Inputs (observation and
    every random draw, tests/synth.py) are generated OUTSIDE the timed region."""
    from oracle import tdmpc_oracle as O
    if kind == "tdmpc":
        st = {"prev": None}

        def make(i):
            return synth.make_obs(cfg, 9000 + i), synth.make_noise(cfg, 9000 + i)

        def run(inp, i):
            _, _, st["prev"], _ = O.plan_tdmpc(sd, cfg, inp[0], eval_mode=True, step=10 ** 6, t0=(i == 0), prev_mean=st["prev"],
                                               std_floor=cfg.min_std, noise=inp[1])
        return make, run
    drnn = kind == "drnn"
    st = {"plan_horizon": 1, "std": cfg.min_std}
    hid = {"h": torch.zeros(1, cfg.hidden_dim, device=device)}

    def make(i):
        have = i > 0
        sizes = O.icem_drnn_sizes(cfg, 0.5, have, i == 0) if drnn else O.icem_sizes(cfg, 0.5, have)
        return synth.make_obs(cfg, 9000 + i), synth.make_icem_noise(cfg, 9000 + i, sizes, cfg.horizon, have and not (drnn and i == 0), single_beta=drnn)

    def run(inp, i):
        if drnn:
            _, hid["h"], _, _ = O.plan_icem_drnn(sd, cfg, inp[0], hid["h"], eval_mode=True, step=10 ** 6, t0=(i == 0), state=st, noise=inp[1])
        else:
            O.plan_icem_mlp(sd, cfg, inp[0], eval_mode=True, step=10 ** 6, t0=(i == 0), state=st, noise=inp[1])
    return make, run


def cpu_plan_rate(cfg, sd, budget_s, max_plans=64, kind="tdmpc", device=None):
    """The reference's plan() restated (oracle/tdmpc_oracle.py issues the same aten ops in the same order as
    /root/reference/src/algorithm/tdmpc.py:94-163 and the iCEM / DRNN planners), timed on this box's host cores -- or, with
    `device`, the same eager op sequence on the GPU (stock aten CUDA kernels: the reference on its intended device).
    Bounded sample: single-environment plans until `budget_s` seconds are spent; inputs are generated outside the timed region."""
    torch.set_num_threads(os.cpu_count() or 1)
    make, run = _oracle_planner(kind, cfg, sd, device)
    times = []
    t_start = time.perf_counter()
    i = 0
    while True:
        inp = make(i)
        if device is not None:
            mv = lambda x: x.to(device) if torch.is_tensor(x) else x                                  # noqa: E731
            inp = (inp[0], {k: ([mv(x) for x in v] if isinstance(v, list) else mv(v)) for k, v in inp[1].items()})
            torch.cuda.synchronize()
        t = time.perf_counter()
        if device is not None:
            with torch.device(device):
                run(inp, i)
            torch.cuda.synchronize()
        else:
            run(inp, i)
        dt = time.perf_counter() - t
        if i >= 2:            # first calls warm MKL / cuBLAS / the allocator
            times.append(dt)
        i += 1
        if (time.perf_counter() - t_start > budget_s and len(times) >= 3) or len(times) >= max_plans:
            break
    med = float(np.median(times))
    return 1.0 / med, len(times), med


def baselines_for(kind, cfg, sd, dev, budget_s, eager=True):
    """cpu_baseline (+ gpu_eager) objects for one configuration."""
    rate, n, med = cpu_plan_rate(cfg, sd, budget_s, kind=kind)
    out = {"cpu_baseline": {"value": rate, "unit": "plans/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{n} single-env plan() calls (median {med * 1e3:.1f} ms/plan), oracle port"}}
    if eager:
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        rate_g, n_g, med_g = cpu_plan_rate(cfg, sd_dev, min(budget_s, 4.0), kind=kind, device=dev)
        out["gpu_eager"] = {"value": rate_g, "unit": "plans/s", "ms_per_plan": med_g * 1e3, "kind": "port on cuda (stock aten / cuBLAS kernels, ~1.5 k launches per plan)",
                            "sample": f"{n_g} single-env plan() calls, the reference's op sequence with every tensor on the GPU"}
    return out


def run_reference(args, cfg, desc, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the Python reference itself cannot
    travel to the GPU box), all host threads, rank 0 only.  A step = a bounded sample of the batch: `per_step` single-environment
    plans (the reference has no batch API, tdmpc.py:110); observations and noise are generated outside the timed region."""
    if rank != 0:
        return
    sd = synth.make_told_weights(cfg, 0)
    F, S, (H, N, P, T) = flops_and_steps(cfg)
    per_step = 16
    make, run = _oracle_planner("tdmpc", cfg, sd)
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    k = 0
    for s_ in range(args.warmup + args.steps):
        inputs = [make(k + j) for j in range(per_step)]
        t = time.perf_counter()
        for j in range(per_step):
            run(inputs[j], k + j)
        if s_ >= args.warmup:
            times.append(time.perf_counter() - t)
        k += per_step
    ms = 1e3 * float(np.mean(times))
    value = per_step / (ms / 1e3)
    B_total = args.batch or WORKLOADS[args.workload][1]
    line = {
        "impl": "reference", "metric": "plans/sec", "value": value, "unit": "plans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(world, 1), B_total, T),
        "sample_steps_per_sec": value * S,
        "cpu_baseline": {"value": value, "unit": "plans/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{per_step} single-env plan() calls per step x {args.steps} steps, sequential (the reference has no batch API, "
                                   "tdmpc.py:110), oracle port (same aten CPU ops as the reference), inputs generated outside the timed region"},
        "e2e": {"value": value, "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="environments in TOTAL, sharded over the GPUs (default: workload's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg_name, def_batch, desc = WORKLOADS[args.workload]
    cfg = synth.make_cfg(cfg_name)

    if args.impl == "reference":
        run_reference(args, cfg, desc, rank, world)
        return

    import torch.distributed as dist
    from tdmpc_b200 import Planner
    from tdmpc_b200.accounting import value_launch_flops
    from tdmpc_b200.parallel import gather_actions, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # BASELINE configs[2]: B_total environments sharded over the ranks (strong scaling).  A workload with fewer environments than
    # ranks (the B = 1 latency configs) does not shard: replicas only, every rank plans the same problem size.
    B_total = args.batch or def_batch
    replicas = B_total < world
    lo, hi = (0, B_total) if replicas else shard_range(B_total, rank, world)
    B = hi - lo
    planned_total = B_total * world if replicas else B_total
    F, S, (H, N, P, T) = flops_and_steps(cfg)
    sd = synth.make_told_weights(cfg, 0)
    pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, precision=args.precision, device=dev)
    pl.sync_weights({k: v.to(dev) for k, v in sd.items()})
    pl.seed = 1234 + rank
    kw = dict(horizon=H, num_samples=N, num_pi_trajs=P, num_elites=cfg.num_elites, iterations=cfg.iterations,
              discount=cfg.discount, temperature=cfg.temperature, momentum=cfg.momentum, min_std=cfg.min_std,
              std_floor=cfg.min_std, eval_mode=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def measure(Bl, total, steps, warmup, with_e2e=True):
        """Device-timed and end-to-end plans over Bl local environments (`total` across the job)."""
        obs_host = torch.from_numpy(synth.make_obs(cfg, 10 + rank, batch=Bl)).pin_memory()      # pinned host buffer (e2e path)
        obs_dev = obs_host.to(dev)                                                              # device-resident copy (kernel path)
        prev_mean = torch.zeros(Bl, H, cfg.action_dim, device=dev)
        out = {"action": torch.empty(Bl, cfg.action_dim, device=dev), "metrics": torch.empty(Bl, 2, device=dev)}
        act_host = torch.empty(Bl, cfg.action_dim).pin_memory()
        met_host = torch.empty(Bl, 2).pin_memory()
        gather = world > 1 and not replicas

        def step_device(first):
            pl.plan(obs_dev, prev_mean=prev_mean, has_prev=not first, t0=first, out=out, **kw)
            if gather:
                gather_actions(out["action"], total)

        def step_e2e(first):
            o = obs_host.to(dev, non_blocking=True)                                  # H2D of this step's inputs
            pl.plan(o, prev_mean=prev_mean, has_prev=not first, t0=first, out=out, **kw)
            if gather:
                gather_actions(out["action"], total)
            act_host.copy_(out["action"], non_blocking=True)                         # D2H of this rank's planned actions
            met_host.copy_(out["metrics"], non_blocking=True)                        # ... and plan_metrics
            torch.cuda.current_stream().synchronize()                                # the env loop needs them on the host

        def timed(fn, steps, warmup):
            for i in range(warmup):
                fn(i == 0)
            barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            l0 = pl.launch_count
            sampler = ClockSampler(local_rank) if rank == 0 else None
            for i in range(steps):
                flush.fill_(i & 0xFF)                       # L2 flush between timed iterations (outside the event pair)
                ev[i][0].record()
                fn(False)
                ev[i][1].record()
            barrier()
            clocks = sampler.stop() if sampler else None
            ms = sum(a.elapsed_time(b) for a, b in ev)
            return max_over_ranks(ms) / steps, pl.launch_count - l0, clocks          # max over ranks

        ms_dev, launches, clocks = timed(step_device, steps, warmup)
        ms_e2e = timed(step_e2e, steps, 1)[0] if with_e2e else None
        return ms_dev, ms_e2e, launches, clocks, obs_dev

    ms_dev, ms_e2e, launches, clocks, obs_dev = measure(B, planned_total, args.steps, args.warmup)

    extra = {}
    if world > 1 and not replicas and not args.no_extra:
        # the weak-scaling line of the same workload: def_batch environments PER GPU
        w_ms, w_e2e, _, _, _ = measure(def_batch, def_batch * world, max(2, min(args.steps, 3)), 3)
        extra["weak_scaling"] = {"envs_per_gpu": def_batch, "envs_total": def_batch * world, "ms_per_step": w_ms,
                                 "plans_per_sec": def_batch * world / (w_ms * 1e-3), "e2e_plans_per_sec": def_batch * world / (w_e2e * 1e-3)}
    if world == 8 and not args.no_extra:
        # BASELINE configs[4]: iCEM through the GRU latent dynamics, 65,536 environments x 1024 samples across 8 B200 (8192 per GPU)
        barrier()
        r5 = drnn_batched(dev, args.precision, 8192, reps=2, roofline=(rank == 0))
        ms5 = max_over_ranks(r5["ms_per_batched_plan"])
        r5.update({"envs_total": 65536, "envs_per_gpu": 8192, "ms_per_batched_plan": ms5, "plans_per_sec": 65536 / ms5 * 1e3,
                   "sample_steps_per_sec": r5["sample_steps_per_sec"] / r5["plans_per_sec"] * 65536 / ms5 * 1e3 if r5["plans_per_sec"] else None,
                   "tflops": 65536 / ms5 * 1e3 * r5["flop_per_plan"] / 1e12})
        extra["cfg5_65536_envs_8gpu"] = r5

    # dominant kernel alone: one estimate_value launch over all B*T local trajectories (same shape the plan uses)
    roof = None
    if rank == 0:
        sus, burst, hbm, src = measured_peaks()
        L, A, M = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim
        f_launch = value_launch_flops(B, T, H, A, L, M)
        z0 = pl.encode(obs_dev)
        acts = torch.rand(B, T, H, A, device=dev) * 2 - 1
        k_ms = time_kernel(lambda: pl.estimate_value(z0, acts, discount=cfg.discount, min_std=cfg.min_std, seed=1, call=1))
        achieved = f_launch / (k_ms * 1e-3) / 1e12
        traffic, traffic_note = None, "no ncu capture of this workload / batch under profiles/"
        tp = os.path.join(ROOT, "profiles", "rollout_tc_traffic.json")
        if os.path.exists(tp):
            t = json.load(open(tp)).get(args.workload)
            if t:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from an ncu capture (per launch)
                traffic = t["dram_bytes_per_launch"] * B / t["batch"]
                traffic_note = t.get("note", "") + ("" if t["batch"] == B else f" (captured at {t['batch']} environments per launch, scaled by {B}/{t['batch']}: DRAM traffic is per-trajectory data only)")
        roof = {"bound": "tensor", "kernel": "rollout_tc_kernel (value mode)" if args.precision == "bf16" else "rollout_fp32_kernel",
                "achieved": achieved, "peak": sus, "unit": "TFLOP/s", "frac": achieved / sus, "traffic": traffic, "traffic_note": traffic_note,
                "algorithmic_bytes": float(B) * T * (H * A * 4 + 4 + 4) + B * L * 4,
                "peak_source": f"{src}: bf16_tflops_sustained (kernel runs back-to-back for >1 s); burst {burst}",
                "kernel_ms": k_ms, "flop_per_launch": f_launch, "envs_per_launch": B,
                "share_of_step": cfg.iterations * k_ms / ms_dev}
        del acts

    if rank == 0 and not args.no_extra:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # B=1 latency cases (BASELINE.json configs[0], configs[1]): plans/s of a single environment, next to the reference's op
        # sequence on the host cores and -- the honest comparator for one environment -- eagerly on this GPU
        for wn in ("cfg1", "cfg2"):
            c2 = synth.make_cfg(WORKLOADS[wn][0])
            f2, s2, (h2, n2, p2, t2) = flops_and_steps(c2)
            sd2 = synth.make_told_weights(c2, 0)
            p1 = Planner(c2.obs_shape[0], c2.action_dim, c2.latent_dim, c2.mlp_dim, c2.enc_dim, precision=args.precision, device=dev)
            p1.sync_weights({k: v.to(dev) for k, v in sd2.items()})
            o1 = torch.from_numpy(synth.make_obs(c2, 3, batch=1)).to(dev)
            pm = torch.zeros(1, h2, c2.action_dim, device=dev)
            k2 = dict(horizon=h2, num_samples=n2, num_pi_trajs=p2, num_elites=c2.num_elites, iterations=c2.iterations,
                      discount=c2.discount, temperature=c2.temperature, momentum=c2.momentum, min_std=c2.min_std,
                      std_floor=c2.min_std, eval_mode=True)
            for i in range(5):
                p1.plan(o1, prev_mean=pm, has_prev=i > 0, t0=i == 0, **k2)
            torch.cuda.synchronize()
            e0.record()
            for i in range(20):
                p1.plan(o1, prev_mean=pm, has_prev=True, t0=False, **k2)
            e1.record(); torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / 20
            z1 = p1.encode(o1)
            a1 = torch.rand(1, t2, h2, c2.action_dim, device=dev) * 2 - 1
            k1 = time_kernel(lambda: p1.estimate_value(z1, a1, discount=c2.discount, min_std=c2.min_std, seed=1, call=1), reps=5)
            fl1 = value_launch_flops(1, t2, h2, c2.action_dim, c2.latent_dim, c2.mlp_dim)
            sus = measured_peaks()[0]
            extra[wn + "_b1"] = {"workload": WORKLOADS[wn][2], "ms_per_plan": ms1, "plans_per_sec": 1e3 / ms1,
                                 "sample_steps_per_sec": 1e3 / ms1 * s2, "flop_per_plan": f2,
                                 "roofline": {"bound": "tensor", "kernel": "rollout_tc_kernel (value mode), one environment = 5 tiles on 5 of 148 SMs",
                                              "achieved": fl1 / (k1 * 1e-3) / 1e12, "peak": sus, "unit": "TFLOP/s", "frac": fl1 / (k1 * 1e-3) / 1e12 / sus,
                                              "traffic": None, "kernel_ms": k1},
                                 **baselines_for("tdmpc", c2, sd2, dev, 4.0)}
        # BASELINE configs[1] dims batched (1024 environments): what the kernel does at dog-run dims when the GPU is full
        c2 = synth.make_cfg("cfg2")
        f2, s2, (h2, n2, p2, t2) = flops_and_steps(c2)
        p1 = Planner(c2.obs_shape[0], c2.action_dim, c2.latent_dim, c2.mlp_dim, c2.enc_dim, precision=args.precision, device=dev)
        p1.sync_weights({k: v.to(dev) for k, v in synth.make_told_weights(c2, 0).items()})
        Bd = 1024
        od = torch.from_numpy(synth.make_obs(c2, 3, batch=Bd)).to(dev)
        pmd = torch.zeros(Bd, h2, c2.action_dim, device=dev)
        k2 = dict(horizon=h2, num_samples=n2, num_pi_trajs=p2, num_elites=c2.num_elites, iterations=c2.iterations, discount=c2.discount,
                  temperature=c2.temperature, momentum=c2.momentum, min_std=c2.min_std, std_floor=c2.min_std, eval_mode=True)
        msd = time_kernel(lambda: p1.plan(od, prev_mean=pmd, has_prev=True, t0=False, **k2), reps=2)
        zd = p1.encode(od)
        ad = torch.rand(Bd, t2, h2, c2.action_dim, device=dev) * 2 - 1
        kd = time_kernel(lambda: p1.estimate_value(zd, ad, discount=c2.discount, min_std=c2.min_std, seed=1, call=1), reps=2)
        fld = value_launch_flops(Bd, t2, h2, c2.action_dim, c2.latent_dim, c2.mlp_dim)
        extra["cfg2_b1024"] = {"workload": WORKLOADS["cfg2"][2] + f", {Bd} environments per launch", "ms_per_batched_plan": msd,
                               "plans_per_sec": Bd / msd * 1e3, "sample_steps_per_sec": Bd / msd * 1e3 * s2, "tflops": Bd / msd * 1e3 * f2 / 1e12,
                               "roofline": {"bound": "tensor", "kernel": "rollout_tc_kernel (value mode)", "achieved": fld / (kd * 1e-3) / 1e12,
                                            "peak": measured_peaks()[0], "unit": "TFLOP/s", "frac": fld / (kd * 1e-3) / 1e12 / measured_peaks()[0],
                                            "traffic": None, "kernel_ms": kd, "flop_per_launch": fld}}
        del ad, od, p1
        # iCEM planners through the drop-in agent classes, B = 1 (BASELINE.json configs[3], configs[4] dims), built-in
        # Philox / colored-noise sampler: both on the tcgen05 path in bf16 mode (GRU: value rollouts; its policy rollout stays fp32)
        from tdmpc_b200 import TdICemSimDssm, TdICemSimMlp
        for wn, cls, prec, mk in (("cfg4", TdICemSimMlp, args.precision, lambda c: synth.make_told_weights(c, 0, enc_ln=True)),
                                  ("cfg5", TdICemSimDssm, args.precision, lambda c: synth.make_drnn_weights(c, 0))):
            c2 = synth.make_cfg(wn, device=str(dev), precision=prec)
            ag = cls(c2)
            sd2 = mk(c2)
            ag.model.load_state_dict(sd2, strict=False)
            ag.model.eval()
            ag.std = c2.min_std
            o1 = synth.make_obs(c2, 3)
            hid = torch.zeros(1, c2.hidden_dim, device=dev)

            def one(i, ag=ag, o1=o1, drnn=(wn == "cfg5")):
                nonlocal hid
                if drnn:
                    _, hid, _ = ag.plan(o1, hid, eval_mode=True, step=10 ** 6, t0=i == 0)
                else:
                    ag.plan(o1, eval_mode=True, step=10 ** 6, t0=i == 0)
            for i in range(4):
                one(i)
            torch.cuda.synchronize()
            t_0 = time.perf_counter()
            for i in range(10):
                one(i + 4)
            torch.cuda.synchronize()
            ms1 = (time.perf_counter() - t_0) * 1e3 / 10
            extra[wn + "_b1"] = {"workload": f"{cls.__name__} (iCEM, {'GRU' if wn == 'cfg5' else 'MLP'} dynamics): obs 18, act 4, H=10, N=1024, K=32, "
                                             f"mixture 0.5, I=6, precision {prec}", "ms_per_plan_wall": ms1, "plans_per_sec": 1e3 / ms1,
                                 **baselines_for("drnn" if wn == "cfg5" else "icem", synth.make_cfg(wn), sd2, dev, 4.0)}
        # BASELINE.json configs[4]: batched iCEM planning through the GRU latent dynamics (256 environments per launch here)
        extra["cfg5_b256"] = drnn_batched(dev, args.precision, 256, roofline=True)
        # BASELINE.json configs[3]: the same batched through the MLP dynamics (TdICemSimMlp)
        extra["cfg4_b256"] = icem_batched(dev, args.precision, 256, False, roofline=True)
        if world == 1:
            extra["replay_sample_1m"] = replay_extra(dev, measured_peaks()[2])
            extra["update_forward_b512"] = update_forward_extra(dev)
            extra["train_step_b512"] = train_step_extra(dev)
        # the batched environment-stepping loop (tdmpc_b200.PlanLoop): numpy observations in, numpy
        # actions out, one pinned H2D + one CUDA-graph replay of the whole plan + one D2H per step (wall clock, host included)
        from tdmpc_b200 import PlanLoop
        for wn, bl in (("cfg3", B), ("cfg1", 1)):
            c2 = synth.make_cfg(WORKLOADS[wn][0])
            _, _, (h2, n2, p2, t2) = flops_and_steps(c2)
            p1 = pl if wn == args.workload else Planner(c2.obs_shape[0], c2.action_dim, c2.latent_dim, c2.mlp_dim, c2.enc_dim,
                                                        precision=args.precision, device=dev)
            if p1 is not pl:
                p1.sync_weights({k: v.to(dev) for k, v in synth.make_told_weights(c2, 0).items()})
            k2 = dict(horizon=h2, num_samples=n2, num_pi_trajs=p2, num_elites=c2.num_elites, iterations=c2.iterations,
                      discount=c2.discount, temperature=c2.temperature, momentum=c2.momentum, min_std=c2.min_std, std_floor=c2.min_std)
            res = {}
            for graph in (True, False):
                loop = PlanLoop(p1, bl, k2, eval_mode=True, use_graph=graph)
                on = synth.make_obs(c2, 5, batch=bl)
                for i in range(3):
                    loop.step(on, i == 0)
                nrep = 5 if bl > 1 else 50
                t_0 = time.perf_counter()
                for i in range(nrep):
                    loop.step(on, False)
                res[graph] = (time.perf_counter() - t_0) * 1e3 / nrep
            extra[f"plan_loop_{wn}_b{bl}"] = {"workload": WORKLOADS[wn][2], "envs": bl, "ms_per_step_graph": res[True], "ms_per_step_eager": res[False],
                                             "plans_per_sec_graph": bl * 1e3 / res[True], "h2d_bytes_per_step": loop.h2d_bytes_per_step,
                                             "d2h_bytes_per_step": loop.d2h_bytes_per_step, "timing": "wall clock incl. host copies"}

    if rank == 0:
        cpu_rate, n_cpu, med = cpu_plan_rate(cfg, sd, args.cpu_budget) if world == 1 else (None, 0, None)
        plans_per_s = planned_total / (ms_dev * 1e-3)
        e2e_rate = planned_total / (ms_e2e * 1e-3)
        line = {
            "metric": "plans/sec", "value": plans_per_s, "unit": "plans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak" if replicas else "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world, B_total, T),
            "sample_steps_per_sec": plans_per_s * S,
            "flop_per_plan": F, "tflops": plans_per_s * F / 1e12,
            "e2e": {"value": e2e_rate, "unit": "plans/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": B * cfg.obs_shape[0] * 4, "d2h_bytes_per_step": B * cfg.action_dim * 4 + B * 2 * 4,
                    "note": "bytes per rank; every rank copies its own shard in and its own actions / metrics out"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "extra": extra,
        }
        if replicas:
            line["config"]["note"] = "fewer environments than ranks: replicas only (a single plan() does not split, SURVEY 8e)"
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "plans/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n_cpu} single-env plan() calls of the same workload (median {med * 1e3:.1f} ms/plan), inputs generated "
                                              f"outside the timed region; the reference has no batch API, so B envs cost B/value seconds"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

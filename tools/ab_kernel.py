"""A/B timing of rollout_tc_kernel builds: kernel ms, SM clock and board power while the kernel runs back to back.
  python tools/ab_kernel.py --lib build/variants/x.so [--batch 512 4096] [--seconds 1.0] [--workload cfg3]
The variant library replaces tdmpc_b200/libtdmpc_b200.so for this process only (experiments; not a product path)."""
import argparse, os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None); ap.add_argument("--batch", type=int, nargs="+", default=[512, 4096])
ap.add_argument("--seconds", type=float, default=1.0); ap.add_argument("--workload", default="cfg3")
a = ap.parse_args()
from tdmpc_b200 import _lib
if a.lib: _lib.LIB_PATH = os.path.abspath(a.lib)
from tests import synth
from tdmpc_b200 import Planner
import pynvml
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
cfg = synth.make_cfg(a.workload)
H, N, P, T = synth.plan_dims(cfg)
drnn = a.workload in ("cfg5", "tiny_drnn")
sd = synth.make_drnn_weights(cfg, 0) if drnn else synth.make_told_weights(cfg, 0)
pl = Planner(cfg.obs_shape[0], cfg.action_dim, cfg.latent_dim, cfg.mlp_dim, cfg.enc_dim, precision="bf16", enc_layernorm=drnn,
             dynamics="drnn" if drnn else "mlp", hidden_dim=cfg.hidden_dim if drnn else 0)
pl.sync_weights({k: v.cuda() for k, v in sd.items()})
for B in a.batch:
    z0 = pl.encode(torch.from_numpy(synth.make_obs(cfg, 1, batch=B)).cuda())
    acts = torch.rand(B, T, H, cfg.action_dim, device="cuda") * 2 - 1
    kw = dict(discount=cfg.discount, min_std=cfg.min_std, seed=1, call=1)
    if drnn: kw["hidden"] = torch.zeros(B, cfg.hidden_dim, device="cuda")
    for _ in range(3): pl.estimate_value(z0, acts, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pl.estimate_value(z0, acts, **kw); e1.record(); torch.cuda.synchronize()
    reps = max(3, int(a.seconds * 1e3 / e0.elapsed_time(e1)))
    samples, stop = [], False
    def poll():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3))
            time.sleep(0.01)
    th = threading.Thread(target=poll); th.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in evs:
        s.record(); pl.estimate_value(z0, acts, **kw); e.record()
    torch.cuda.synchronize(); stop = True; th.join()
    ms = np.array([s.elapsed_time(e) for s, e in evs])
    half = samples[len(samples) // 2:] or samples
    print(f"{os.path.basename(a.lib or 'default')}: B={B} reps={reps} ms first={ms[0]:.3f} median={np.median(ms):.3f} last={ms[-1]:.3f} "
          f"min={ms.min():.3f} | clk {np.median([c for c, _ in half]):.0f} MHz  power {np.median([p for _, p in half]):.0f} W", flush=True)

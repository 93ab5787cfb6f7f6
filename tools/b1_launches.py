"""One-environment plan() latency by kernel (BASELINE configs[0] / configs[1]):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b1_cfg2.csv python tools/b1_launches.py cfg2"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth  # noqa: E402
from tdmpc_b200 import Planner  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
c = synth.make_cfg(name)
H, N, P, T = synth.plan_dims(c)
pl = Planner(c.obs_shape[0], c.action_dim, c.latent_dim, c.mlp_dim, c.enc_dim, precision="bf16")
pl.sync_weights({k: v.cuda() for k, v in synth.make_told_weights(c, 0).items()})
o = torch.from_numpy(synth.make_obs(c, 3, batch=1)).cuda()
pm = torch.zeros(1, H, c.action_dim, device="cuda")
kw = dict(horizon=H, num_samples=N, num_pi_trajs=P, num_elites=c.num_elites, iterations=c.iterations, discount=c.discount,
          temperature=c.temperature, momentum=c.momentum, min_std=c.min_std, std_floor=c.min_std, eval_mode=True)
for i in range(3):
    pl.plan(o, prev_mean=pm, has_prev=i > 0, t0=i == 0, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    pl.plan(o, prev_mean=pm, has_prev=True, t0=False, **kw)
e1.record(); torch.cuda.synchronize()
print(name, "ms per plan", e0.elapsed_time(e1) / 10, "launches per plan", pl.launch_count / 13)

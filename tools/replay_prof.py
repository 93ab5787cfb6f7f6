"""Drives tdmpc_b200.ReplayBuffer.sample at the reference's full size for ncu (launch list / full set of the replay kernels).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/replay_launches.csv python tools/replay_prof.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth  # noqa: E402
from tdmpc_b200 import ReplayBuffer  # noqa: E402

full = len(sys.argv) > 1 and sys.argv[1] == "full"
cfg = synth.make_replay_cfg("small", train_steps=1000000, episode_length=1000, batch_size=512, obs_shape=(18,), action_dim=4)
buf = ReplayBuffer(cfg, latent_plan=True)
rs = np.random.RandomState(7)
pri = np.exp(1.5 * rs.randn(buf.capacity)).astype(np.float32)
pri[(np.arange(buf.capacity) % 1000) >= 995] = 0
buf._priorities.copy_(torch.from_numpy(pri))
buf.idx, buf._full = (0, True) if full else (buf.capacity - 1000, False)
for _ in range(3):
    out = buf.sample()
torch.cuda.synchronize()
print("rounds", buf.last_rounds, "idx[:4]", out[4][:4].tolist())

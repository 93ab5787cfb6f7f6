"""Device time of select_kernel alone at small batches (select + distribution update, no sampling): back-to-back calls, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tdmpc_b200 import Planner  # noqa: E402

pl = Planner(18, 4, 50, precision="fp32")
for B, T, K, H, A in ((1, 537, 64, 5, 1), (1, 537, 64, 5, 38), (1, 1544, 32, 10, 4), (4096, 537, 64, 5, 4)):
    pl.lib.tdmpc_reserve(pl._h, B, T, H, K)
    v = torch.randn(B, T, device="cuda"); a = torch.rand(B, T, H, A, device="cuda") * 2 - 1
    mean = torch.zeros(B, H, A, device="cuda"); std = torch.ones(B, H, A, device="cuda")
    kw = dict(num_elites=K, temperature=0.5, momentum=0.1, std_floor=0.05)
    for _ in range(5):
        pl.select_update(v, a, mean, std, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if B == 1 else 20
    e0.record()
    for _ in range(n):
        pl.select_update(v, a, mean, std, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"select_update B={B} T={T} K={K} H={H} A={A}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call")

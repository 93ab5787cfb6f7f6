"""Time of one whole training step (tdmpc_update: forward, backward, clip, Adam x2, update_pi, ema) on the device against the same
step done by torch eager on this GPU with the oracle's op sequence (autograd + torch.optim.Adam) and on the host cores.
  python tools/train_bench.py [--batch 512] [--workload cfg3] [--reps 20]"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import synth
from tdmpc_b200 import TDMPC
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=512); ap.add_argument("--workload", default="cfg3")
ap.add_argument("--reps", type=int, default=20); ap.add_argument("--eager", type=int, default=1)
a = ap.parse_args()
cfg = synth.make_cfg(a.workload, batch_size=a.batch, device="cuda", precision="fp32")
sd, sd_t = synth.make_told_weights(cfg, 0), synth.make_told_weights(cfg, 50)
agent = TDMPC(cfg); agent.model.load_state_dict(sd); agent.model_target.load_state_dict(sd_t)
b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_train_batch(cfg, 1, a.batch).items()}
for _ in range(3): agent.update_batch(b["obs"], b["next_obses"], b["action"], b["reward"], b["weights"], 0)
torch.cuda.synchronize()
n0 = agent._planner.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.reps): agent.update_batch(b["obs"], b["next_obses"], b["action"], b["reward"], b["weights"], i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"tdmpc_update {a.workload} batch {a.batch}: {ms:.3f} ms per step, {(agent._planner.launch_count - n0) // a.reps} launches per step")
if a.eager:
    from oracle import tdmpc_oracle as O                      # bench comparator only (the reference's op sequence, eager aten / cuBLAS kernels)
    osd, ot, opt = {k: v.cuda() for k, v in sd.items()}, {k: v.cuda() for k, v in sd_t.items()}, None
    for _ in range(2): _, osd, ot, _, opt = O.update_step(osd, ot, cfg, b, opt, True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): _, osd, ot, _, opt = O.update_step(osd, ot, cfg, b, opt, True)
    torch.cuda.synchronize(); print(f"torch eager on this GPU (oracle op sequence): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per step")
    bc = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in b.items()}
    t0 = time.perf_counter(); O.update_step(sd, sd_t, cfg, bc, None, True)
    print(f"host cores ({torch.get_num_threads()} threads): {(time.perf_counter() - t0) * 1e3:.1f} ms per step")

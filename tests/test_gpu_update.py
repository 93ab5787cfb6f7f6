"""Forward half of the training step through the C ABI (tdmpc_update_forward, fp32) against the oracle
and against what the unmodified reference TDMPC.update returned (tests/golden/update_*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import tdmpc_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu


def _agent(cfg, sd, sd_t):
    from tdmpc_b200 import TDMPC
    agent = TDMPC(cfg)
    agent.model.load_state_dict(sd)
    agent.model_target.load_state_dict(sd_t)
    return agent


@pytest.mark.parametrize("name,cfg_name", [("update_tiny", "tiny"), ("update_cfg3", "cfg3"), ("update_cfg2s", "cfg2s")])
def test_update_forward_matches_oracle_and_reference_golden(golden_dir, name, cfg_name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    seed, batch = [int(x) for x in g["meta"]]
    cfg = synth.make_cfg(cfg_name, batch_size=batch, device="cuda", precision="fp32")
    sd, sd_t = synth.make_told_weights(cfg, seed), synth.make_told_weights(cfg, seed + 50)
    b = synth.make_update_batch(cfg, seed, batch)
    agent = _agent(cfg, sd, sd_t)
    r = agent.update_forward(b["obs"], b["next_obses"], b["action"], b["reward"], weights=b["weights"], noise_pi=b["noise_pi"], taps=True)
    torch.cuda.synchronize()
    o = O.update_forward(sd, sd_t, cfg, b)
    # per-step taps, then the per-sample losses: 1e-5 relative (fp32 mode)
    for k in ("td_target", "q1", "q2", "reward_pred"):
        scale = o["taps"][k].abs().max().item()
        assert (r["taps"][k].cpu() - o["taps"][k]).abs().max().item() <= 1e-5 * max(scale, 1.0), k
    np.testing.assert_allclose(r["zs"].cpu().numpy(), o["zs"].numpy(), rtol=0, atol=2e-5)
    for k in ("consistency", "reward", "value", "priority", "total"):
        np.testing.assert_allclose(r["per_sample"][k].cpu().numpy(), o[k].squeeze(1).numpy(), rtol=2e-5, atol=2e-5, err_msg=k)
    # the reference's own numbers
    got = [r["consistency_loss"], r["reward_loss"], r["value_loss"], r["total_loss"], r["weighted_loss"]]
    np.testing.assert_allclose(got, g["metrics"][:5], rtol=2e-5)
    np.testing.assert_allclose(r["priorities"].cpu().numpy(), g["priorities"], rtol=2e-5, atol=2e-5)


def test_update_forward_uses_both_models_and_follows_weight_updates():
    """The TD target mixes the two parameter sets (online encoder + policy, target Q heads; next_z from the target encoder):
    swapping them must change the result, an in-place update of the target (h.ema, tdmpc.py:237) must be picked up, and the
    built-in Philox draws must replay per (seed, call)."""
    cfg = synth.make_cfg("tiny", batch_size=40, device="cuda", precision="fp32")
    sd, sd_t = synth.make_told_weights(cfg, 1), synth.make_told_weights(cfg, 51)
    b = synth.make_update_batch(cfg, 3, 40)
    agent = _agent(cfg, sd, sd_t)
    kw = dict(weights=b["weights"], noise_pi=b["noise_pi"])
    r0 = agent.update_forward(b["obs"], b["next_obses"], b["action"], b["reward"], **kw)
    swapped = _agent(cfg, sd_t, sd).update_forward(b["obs"], b["next_obses"], b["action"], b["reward"], **kw)
    assert abs(r0["value_loss"] - swapped["value_loss"]) > 1e-3 * abs(r0["value_loss"])
    with torch.no_grad():                                       # ema: p_target.lerp_(p, tau)
        for p, pt in zip(agent.model.parameters(), agent.model_target.parameters()):
            pt.data.lerp_(p.data, 0.5)
    sd_t2 = {k: v.detach().cpu() for k, v in agent.model_target.state_dict().items()}
    r1 = agent.update_forward(b["obs"], b["next_obses"], b["action"], b["reward"], **kw)
    o1 = O.update_forward(sd, sd_t2, cfg, b)
    np.testing.assert_allclose(r1["per_sample"]["total"].cpu().numpy(), o1["total"].squeeze(1).numpy(), rtol=2e-5, atol=2e-5)
    # without injected draws: Philox keyed by (seed, call) -- same key, same result; next call, fresh draws
    agent._planner.call = 7
    a = agent.update_forward(b["obs"], b["next_obses"], b["action"], b["reward"])
    agent._planner.call = 7
    a2 = agent.update_forward(b["obs"], b["next_obses"], b["action"], b["reward"])
    c = agent.update_forward(b["obs"], b["next_obses"], b["action"], b["reward"])
    assert torch.equal(a["per_sample"]["value"], a2["per_sample"]["value"]) and not torch.equal(a["per_sample"]["value"], c["per_sample"]["value"])

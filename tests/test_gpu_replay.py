"""GPU parity of the prioritised replay sampler (tdmpc_b200.ReplayBuffer -> tdmpc_replay_sample / tdmpc_replay_gather)
against the reference's own outputs (tests/golden/replay_*.npz) and the numpy oracle.  Index work is bit-exact;
the importance weights (fp32 pow) agree to 2e-6 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import replay_oracle as RO
from tests import synth

pytestmark = pytest.mark.gpu


class Ep:   # what ReplayBuffer.add reads from a helper.Episode
    def __init__(self, obs, action, reward):
        self.obs, self.action, self.reward = torch.from_numpy(obs), torch.from_numpy(action), torch.from_numpy(reward)


def filled(cfg, seed):
    from tdmpc_b200 import ReplayBuffer
    buf = ReplayBuffer(cfg, latent_plan=True)
    for obs, act, rew in synth.make_replay_episodes(cfg, seed):
        buf += Ep(obs, act, rew)
    n_valid = buf.capacity if buf._full else buf.idx
    for idxs, pr in synth.make_priority_updates(cfg, n_valid, seed):
        buf.update_priorities(torch.from_numpy(idxs).cuda(), torch.from_numpy(pr).cuda())
    return buf


@pytest.mark.parametrize("name,cfg_name,seed", [("replay_small", "small", 0), ("replay_full", "full", 1)])
def test_replay_sample_matches_reference_golden(golden_dir, name, cfg_name, seed):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    cfg = synth.make_replay_cfg(cfg_name)
    buf = filled(cfg, seed)
    assert np.array_equal(buf._priorities.cpu().numpy(), g["priorities"])           # add / update_priorities mirror
    assert [buf.idx, int(buf._full)] == g["idx_full"].tolist()
    uni, used = synth.make_uniforms(cfg.batch_size * 16 * int(g["meta"][1]), seed), 0
    for d in range(int(g["meta"][1])):
        obs, next_obs, action, reward, idxs, weights = buf.sample(uniforms=uni[used:])
        used += int(g[f"d{d}_used"][0])
        assert np.array_equal(idxs.cpu().numpy(), g[f"d{d}_idxs"]), f"draw {d}: indices differ from the reference"
        for t, key in ((obs, "obs"), (next_obs, "next_obs"), (action, "action"), (reward, "reward")):
            assert t.shape == g[f"d{d}_{key}"].shape and np.array_equal(t.cpu().numpy(), g[f"d{d}_{key}"]), key
        np.testing.assert_allclose(weights.cpu().numpy(), g[f"d{d}_weights"], rtol=2e-6, atol=0)
        if cfg_name == "full":
            assert buf.last_rounds > 1          # numpy's redraw loop really ran on the device


@pytest.mark.parametrize("replace", [True, False])
def test_replay_million_priorities_match_numpy_choice(replace):
    """Full size (cfg.max_buffer_size = 1e6, batch 512): same uniforms -> same indices as numpy's algorithm."""
    from tdmpc_b200 import ReplayBuffer
    cfg = synth.make_replay_cfg("small", train_steps=1000000, episode_length=1000, batch_size=512, obs_shape=(18,), action_dim=4)
    buf = ReplayBuffer(cfg, latent_plan=True)
    rs = np.random.RandomState(7)
    pri = np.exp(1.5 * rs.randn(buf.capacity)).astype(np.float32)
    pri[(np.arange(buf.capacity) % 1000) >= 995] = 0
    buf._priorities.copy_(torch.from_numpy(pri))
    buf._obs.copy_(torch.arange(buf.capacity + 1, dtype=torch.float32, device="cuda")[:, None].expand(-1, 18))
    buf.idx, buf._full = (0, True) if not replace else (buf.capacity - 1000, False)
    total = buf.capacity if buf._full else buf.idx
    uni = synth.make_uniforms(512 * 16, 3)
    idxs, weights = buf.sample_indices(uniforms=uni)
    p = torch.from_numpy(pri[:total]) ** cfg.per_alpha
    p /= p.sum()
    want, _ = RO.choice(total, 512, p.numpy(), replace, uni)
    got = idxs.cpu().numpy()
    # the device carries probs.sum() in float64 (torch's fp32 reduction order is unspecified), so a draw that lands within
    # fp32 rounding of a cdf step may legitimately differ by one slot; everything else is exact
    cdf = np.cumsum(p.numpy().astype(np.float64)); cdf /= cdf[-1]
    bad = [j for j in np.nonzero(want != got)[0] if replace and abs(cdf[min(want[j], got[j])] - uni[j]) > 1e-6]
    assert (want != got).sum() <= 1 and not bad, (np.nonzero(want != got)[0][:8], want[want != got][:8], got[want != got][:8])
    w_ref = (total * p[torch.from_numpy(got)]) ** (-cfg.per_beta)
    np.testing.assert_allclose(weights.cpu().numpy(), (w_ref / w_ref.max()).numpy(), rtol=5e-6)
    assert (pri[got] > 0).all() and (replace or len(set(got.tolist())) == 512)
    obs, next_obs, action, reward = buf.gather(idxs)                                   # windows start at the drawn rows
    assert torch.equal(obs[:, 0].cpu(), idxs.cpu().float()) and torch.equal(next_obs[2, :, 0].cpu(), (idxs + 3).cpu().float())
    again, _ = buf.sample_indices(uniforms=uni)
    assert torch.equal(again, idxs)


def test_replay_builtin_stream_frequencies_and_replay():
    """Philox stream: draws follow priorities ** alpha, and (seed, call) replays bit for bit."""
    from tdmpc_b200 import ReplayBuffer
    cfg = synth.make_replay_cfg("small", batch_size=4096, train_steps=2000)
    a, b = ReplayBuffer(cfg, latent_plan=True, seed=11), ReplayBuffer(cfg, latent_plan=True, seed=11)
    pri = torch.zeros(2000)
    pri[:16] = torch.arange(1, 17).float()
    for buf in (a, b):
        buf._priorities.copy_(pri)
        buf.idx = 1500
    counts = torch.zeros(16, dtype=torch.float64)
    for _ in range(8):
        ia, _ = a.sample_indices()
        ib, wb = b.sample_indices()
        assert torch.equal(ia, ib) and int(ia.max()) < 16
        counts += torch.bincount(ia.cpu(), minlength=16).double()
    p = (pri[:16] ** cfg.per_alpha).double()
    p /= p.sum()
    n = counts.sum()
    chi2 = float((((counts - n * p) ** 2) / (n * p)).sum())
    assert chi2 < 50.0, chi2                     # 15 degrees of freedom: P(chi2 > 50) ~ 1e-5
    assert float(wb.max()) == 1.0 and float(wb.min()) > 0.0


def test_replay_edge_cases():
    from tdmpc_b200 import ReplayBuffer
    cfg = synth.make_replay_cfg("small", batch_size=1, train_steps=400)
    buf = ReplayBuffer(cfg, latent_plan=True)
    with pytest.raises(RuntimeError):
        buf.sample()                                            # empty buffer
    obs, act, rew = synth.make_replay_episodes(cfg, 0)[0]
    buf += Ep(obs, act, rew)
    out = buf.sample(uniforms=np.array([0.999999]))             # batch 1, u close to 1 -> the last drawable transition
    assert int(out[4][0]) == cfg.episode_length - cfg.horizon - 1 and float(out[5][0]) == 1.0
    assert torch.equal(out[1][-1, 0].cpu(), torch.from_numpy(obs[-1]))        # episode end: next_obs[-1] is the stored final observation
    out0 = buf.sample(uniforms=np.array([0.0]))
    assert int(out0[4][0]) == 0 and torch.equal(out0[0][0].cpu(), torch.from_numpy(obs[0]))
    # without replacement there must be batch_size drawable transitions
    cfg2 = synth.make_replay_cfg("small", batch_size=128, train_steps=100, episode_length=100)
    full = ReplayBuffer(cfg2, latent_plan=True)
    full += Ep(obs, act, rew)
    assert full._full
    full._priorities[10:] = 0
    with pytest.raises(RuntimeError):
        full.sample()

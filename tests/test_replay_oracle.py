"""CPU checks of the replay-sampler oracle (oracle/replay_oracle.py): it is numpy's own `choice` given the same
uniform stream, and it reproduces the outputs the reference's ReplayBuffer.sample gave (tests/golden/replay_*.npz,
written by tests/golden/make_golden.py from the unmodified reference, helper.py:435-534)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import replay_oracle as RO
from tests import synth


@pytest.mark.parametrize("replace", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_choice_is_numpy_choice(replace, seed):
    rs = np.random.RandomState(100 + seed)
    n, size = 500, 200
    p = np.exp(2.5 * rs.randn(n)).astype(np.float32)
    p[rs.randint(0, n, 50)] = 0
    p /= p.sum()
    want = np.random.RandomState(seed).choice(n, size, p=p, replace=replace)
    got, used = RO.choice(n, size, p, replace, np.random.RandomState(seed).random_sample(size * 16))
    assert np.array_equal(want, got)
    assert used == size if replace else used >= size
    if not replace:
        assert len(set(got.tolist())) == size


def build_state(cfg, seed):
    st = RO.ReplayState(cfg, latent_plan=True)
    for obs, act, rew in synth.make_replay_episodes(cfg, seed):
        st.add(torch.from_numpy(obs), torch.from_numpy(act), torch.from_numpy(rew))
    n_valid = st["capacity"] if st["full"] else st["idx"]
    for idxs, pr in synth.make_priority_updates(cfg, n_valid, seed):
        st.update_priorities(torch.from_numpy(idxs), torch.from_numpy(pr))
    return st


@pytest.mark.parametrize("name,cfg_name,seed", [("replay_small", "small", 0), ("replay_full", "full", 1)])
def test_oracle_matches_reference_fixture(golden_dir, name, cfg_name, seed):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    cfg = synth.make_replay_cfg(cfg_name)
    st = build_state(cfg, seed)
    assert np.array_equal(st["priorities"].numpy(), g["priorities"])
    assert [st["idx"], int(st["full"])] == g["idx_full"].tolist()
    uni, used = synth.make_uniforms(cfg.batch_size * 16 * int(g["meta"][1]), seed), 0
    for d in range(int(g["meta"][1])):
        out = RO.sample(st, uni[used:])
        used += out[6]
        assert out[6] == int(g[f"d{d}_used"][0])
        assert np.array_equal(out[4].numpy(), g[f"d{d}_idxs"])
        for k, key in enumerate(("obs", "next_obs", "action", "reward")):
            assert np.array_equal(out[k].numpy(), g[f"d{d}_{key}"]), key
        np.testing.assert_allclose(out[5].numpy(), g[f"d{d}_weights"], rtol=1e-6, atol=0)
    if cfg_name == "full":      # the skewed priorities force numpy's redraw rounds
        assert int(g["d0_used"][0]) > cfg.batch_size


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_replay_has_no_cpu_fallback():
    from tdmpc_b200 import ReplayBuffer, _lib
    with pytest.raises(RuntimeError):
        ReplayBuffer(synth.make_replay_cfg("small"))
    lib = _lib.load()
    a = _lib.ReplaySampleArgs()
    assert lib.tdmpc_replay_sample(C.byref(a), None) == -1 and b"null device pointer" in lib.tdmpc_last_error(None)
    assert lib.tdmpc_replay_workspace_bytes(1000000, 512) >= 1000000 + 2 * 8 * 977

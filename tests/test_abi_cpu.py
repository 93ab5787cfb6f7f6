"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
and fails loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest
import torch

from tdmpc_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "tdmpc_b200.h")).read()
    declared = re.findall(r"TDMPC_API\s+[\w\s\*]+?\b(tdmpc_\w+)\s*\(", hdr)
    assert len(declared) >= 15 and set(declared) == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.tdmpc_build_info()


def test_struct_layouts_match_header_sizes():
    # pointer-heavy structs: sizes must be multiples of 8 and the iteration arrays 16 long
    assert C.sizeof(_lib.ModelCfg) == 40
    assert C.sizeof(_lib.Weights) == 8 * (2 + 2 + 2 + 3 * 6 + 6 + 6 + 4 + 4 + 2 + 6 + 4 + 4)
    # the library reports its own sizeof/offsetof values; load() compares them with the ctypes mirror
    got = (C.c_int32 * 11)()
    assert _lib.load().tdmpc_abi_sizes(got, 11) == 11
    assert list(got)[7:] == [C.sizeof(_lib.ReplaySampleArgs), C.sizeof(_lib.ReplayGatherArgs), C.sizeof(_lib.UpdateArgs),
                             C.sizeof(_lib.TrainArgs)]
    assert list(got)[:3] == [C.sizeof(_lib.ModelCfg), C.sizeof(_lib.Weights), C.sizeof(_lib.PlanArgs)]
    assert got[6] == _lib.PlanArgs.tap_z0.offset
    assert _lib.PlanArgs.n_samples.size == 4 * _lib.TDMPC_MAX_ITERS
    assert C.sizeof(_lib.PlanArgs) % 8 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.ModelCfg(5, 1, 50, 512, 256, 128, 0, 0, 0, 0)
    assert lib.tdmpc_create(C.byref(cfg), C.byref(h)) == -3
    assert b"no CPU fallback" in lib.tdmpc_last_error(None)
    from tdmpc_b200 import Planner
    with pytest.raises(RuntimeError):
        Planner(5, 1, 50)


def test_told_state_dict_keys_match_reference_layout():
    from tdmpc_b200 import TOLD
    from tests import synth
    cfg = synth.make_cfg("tiny")
    m = TOLD(cfg)
    assert set(m.state_dict().keys()) == set(synth.make_told_weights(cfg).keys())
    # the reference zero-initialises the reward / Q heads (tdmpc.py:21-23)
    assert float(m._reward[-1].weight.abs().sum()) == 0.0 and float(m._Q2[-1].bias.abs().sum()) == 0.0
    m.load_state_dict(synth.make_told_weights(cfg))


def test_batched_episodes_mirror_reference_episode_layout():
    """tdmpc_b200.BatchedEpisodes keeps, per environment, the fields of the reference's Episode (helper.py:383-415):
    obs [T+1, D] with the reset observation first, action [T, A], reward [T], cumulative_reward, done, first."""
    import numpy as np
    from tdmpc_b200.collector import BatchedEpisodes
    B, T, D, A = 3, 4, 5, 2
    rng = np.random.RandomState(0)
    obs0 = rng.randn(B, D).astype(np.float32)
    ep = BatchedEpisodes(B, T, D, A, obs0)
    assert ep.first and len(ep) == 0
    obs, act, rew = [], [], []
    for t in range(T):
        o, a, r = rng.randn(B, D).astype(np.float32), rng.randn(B, A).astype(np.float32), rng.randn(B).astype(np.float32)
        ep += (o, a, r, t == T - 1)
        obs.append(o); act.append(a); rew.append(r)
    assert len(ep) == T and not ep.first and ep.done.all()
    e1 = ep.episode(1)
    assert len(e1) == T and not e1.first
    assert tuple(e1.obs.shape) == (T + 1, D) and tuple(e1.action.shape) == (T, A) and tuple(e1.reward.shape) == (T,)
    assert np.array_equal(e1.obs[0].numpy(), obs0[1]) and np.array_equal(e1.obs[3].numpy(), obs[2][1])
    assert np.array_equal(e1.action[2].numpy(), act[2][1]) and np.array_equal(e1.reward.numpy(), np.stack(rew)[:, 1])
    assert abs(e1.cumulative_reward - float(np.stack(rew)[:, 1].astype(np.float64).sum())) < 1e-6 and e1.done
    import pytest
    with pytest.raises(IndexError):
        ep.add(obs[0], act[0], rew[0], True)

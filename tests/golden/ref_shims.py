"""Import shims that let the UNMODIFIED reference (/root/reference) run on a CPU-only
container.  Used ONLY by tests/golden/make_golden.py (the fixture generator) -- nothing
under tests/ that runs on the GPU box imports this, because /root/reference does not
exist there.

Shims:
  1. helper.py:12 imports an un-vendored rlpyt fork -> stub modules.
  2. tdmpc.py:60 calls .cuda() -> identity on a CPU-only box.
  3. iCEM/DRNN files import colorednoise, rlpyt utils, gym RunningMeanStd -> stubs.
"""
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference"


def install():
    for p in (REF, REF + "/src"):
        if p not in sys.path:
            sys.path.insert(0, p)

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Dummy:  # placeholder for pixel encoder / optim factory; never used on the state path
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed symbol used")

    _mod("rlpyt"); _mod("rlpyt.ul"); _mod("rlpyt.ul.models"); _mod("rlpyt.ul.models.ul")
    _mod("rlpyt.ul.models.ul.encoders", DmlabEncoderModelNorm=_Dummy)
    _mod("rlpyt.utils")
    _mod("rlpyt.utils.tensor", infer_leading_dims=_Dummy, restore_leading_dims=_Dummy)
    _mod("rlpyt.ul.algos"); _mod("rlpyt.ul.algos.utils")
    _mod("rlpyt.ul.algos.utils.optim_factory", create_optimizer=lambda *a, **k: None)
    _mod("rlpyt.ul.algos.utils.scheduler_factory", create_scheduler=lambda *a, **k: (None, None))     # tdmpc_similarity.py:8

    class RunningMeanStd:
        def __init__(self, *a, **k):
            self.mean, self.var, self.count = 0.0, 1.0, 1e-4

    _mod("gym"); _mod("gym.wrappers")
    _mod("gym.wrappers.normalize", RunningMeanStd=RunningMeanStd)

    if "colorednoise" not in sys.modules:
        import os
        sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
        from oracle.colored_noise import powerlaw_psd_gaussian
        _mod("colorednoise", powerlaw_psd_gaussian=powerlaw_psd_gaussian)

    if not torch.cuda.is_available():
        nn.Module.cuda = lambda self, *a, **k: self

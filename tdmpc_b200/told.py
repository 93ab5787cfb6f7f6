"""Host-side mirror of the reference's TOLD module and schedule helper.

`TOLD` has the same constructor argument, sub-module names, state_dict keys and method surface
(h / next / pi / Q / track_q_grad) as the reference's class (src/algorithm/tdmpc.py:9-50), so a
checkpoint saved by the reference loads here and vice versa.  It stays an ordinary autograd
nn.Module -- training (update / update_pi) is out of scope and keeps running on it unchanged;
the CUDA planner only READS its parameters (tdmpc_b200/planner.py).
"""
import re

import numpy as np
import torch
import torch.nn as nn


def linear_schedule(schdl, step):
    """Same contract as helper.linear_schedule (src/algorithm/helper.py:639-652):
    a float string, or 'linear(init, final, duration, start)'."""
    try:
        return float(schdl)
    except ValueError:
        pass
    m = re.match(r"linear\((.+),(.+),(.+),(.+)\)", schdl)
    if not m:
        raise NotImplementedError(schdl)
    init, final, duration, start = (float(g) for g in m.groups())
    mix = float(np.clip((step - start) / duration, 0.0, 1.0))
    return (1.0 - mix) * init + mix * final


def _mlp(i, h, o):
    return nn.Sequential(nn.Linear(i, h), nn.ELU(), nn.Linear(h, h), nn.ELU(), nn.Linear(h, o))


def _q(i, h):
    return nn.Sequential(nn.Linear(i, h), nn.LayerNorm(h), nn.Tanh(), nn.Linear(h, h), nn.LayerNorm(h), nn.ELU(),
                         nn.Linear(h, 1))


class TruncatedNormalSample:
    """TruncatedNormal(mu, std).sample(clip) of helper.py:71-96, sampling only."""

    @staticmethod
    def sample(mu, std, clip=0.3, eps=1e-6):
        noise = torch.randn_like(mu) * std
        noise = torch.clamp(noise, -clip, clip)
        return torch.clamp(mu + noise, -1.0 + eps, 1.0 - eps)


class TOLD(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        L, A, M, E, D = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim, cfg.enc_dim, cfg.obs_shape[0]
        if getattr(cfg, "modality", "state") != "state":
            raise NotImplementedError("tdmpc_b200 covers the state-observation encoder only (SURVEY.md section 2, row 2)")
        if getattr(cfg, "enc_layernorm", False):
            self._encoder = nn.Sequential(nn.Linear(D, E), nn.LayerNorm(E), nn.ELU(), nn.Linear(E, L))
        else:
            self._encoder = nn.Sequential(nn.Linear(D, E), nn.ELU(), nn.Linear(E, L))
        self._dynamics = _mlp(L + A, M, L)
        self._reward = _mlp(L + A, M, 1)
        self._pi = _mlp(L, M, A)
        self._Q1, self._Q2 = _q(L + A, M), _q(L + A, M)
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.orthogonal_(mod.weight.data)
                nn.init.zeros_(mod.bias)
        for mod in (self._reward, self._Q1, self._Q2):
            mod[-1].weight.data.fill_(0)
            mod[-1].bias.data.fill_(0)

    def track_q_grad(self, enable=True):
        for mod in (self._Q1, self._Q2):
            for p in mod.parameters():
                p.requires_grad_(enable)

    def h(self, obs):
        return self._encoder(obs)

    def next(self, z, a):
        x = torch.cat([z, a], dim=-1)
        return self._dynamics(x), self._reward(x)

    def pi(self, z, std=0):
        mu = torch.tanh(self._pi(z))
        if std > 0:
            return TruncatedNormalSample.sample(mu, torch.ones_like(mu) * std)
        return mu

    def Q(self, z, a):
        x = torch.cat([z, a], dim=-1)
        return self._Q1(x), self._Q2(x)


class DSSM(TOLD):
    """Model of the iCEM agents (src/algorithm/tdmpc_icem_similarity_mlp.py:14-74): TOLD plus a `_predictor`
    head (training only) and, when cfg.normalize, the LayerNorm state encoder (helper.dmlab_enc_norm).
    Same state_dict keys as the reference."""

    def __init__(self, cfg):
        nn.Module.__init__(self)
        self.cfg = cfg
        L, A, M, E, D = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim, cfg.enc_dim, cfg.obs_shape[0]
        if getattr(cfg, "modality", "state") != "state":
            raise NotImplementedError("tdmpc_b200 covers the state-observation encoder only")
        self._dynamics = _mlp(L + A, M, L)
        self._reward = _mlp(L + A, M, 1)
        self._pi = _mlp(L, M, A)
        self._Q1, self._Q2 = _q(L + A, M), _q(L + A, M)
        if cfg.normalize:
            self._encoder = nn.Sequential(nn.Linear(D, E), nn.LayerNorm(E), nn.ELU(), nn.Linear(E, L))
            self._predictor = nn.Sequential(nn.Linear(L, M), nn.BatchNorm1d(M), nn.ELU(), nn.Linear(M, L))
        else:
            self._encoder = nn.Sequential(nn.Linear(D, E), nn.ELU(), nn.Linear(E, L))
            self._predictor = _mlp(L, M, L)
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.orthogonal_(mod.weight.data)
                nn.init.zeros_(mod.bias)
        for mod in (self._reward, self._Q1, self._Q2):
            mod[-1].weight.data.fill_(0)
            mod[-1].bias.data.fill_(0)

    def pred_z(self, z):
        return self._predictor(z)

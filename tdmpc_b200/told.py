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


class NormGRUCell(nn.Module):
    """Bias-free GRU cell with one LayerNorm (eps 1e-3) per gate (src/models/rnns.py:8-29).  Gate chunks along
    3*hidden are reset, update, newval and h' = update * newval + (1 - update) * h."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.weight_ih = nn.Linear(input_size, 3 * hidden_size, bias=False)
        self.weight_hh = nn.Linear(hidden_size, 3 * hidden_size, bias=False)
        self.ln_reset = nn.LayerNorm(hidden_size, eps=1e-3)
        self.ln_update = nn.LayerNorm(hidden_size, eps=1e-3)
        self.ln_newval = nn.LayerNorm(hidden_size, eps=1e-3)

    def forward(self, x, state):
        ri, ui, ni = self.weight_ih(x).chunk(3, 1)
        rh, uh, nh = self.weight_hh(state).chunk(3, 1)
        reset = torch.sigmoid(self.ln_reset(ri + rh))
        update = torch.sigmoid(self.ln_update(ui + uh))
        newval = torch.tanh(self.ln_newval(ni + reset * nh))
        return update * newval + (1 - update) * state


class DGruDyna(nn.Module):
    """Recurrent latent dynamics (src/models/gru_dyna.py:11-29): NormGRUCell([z,a], h) -> prior MLP with
    BatchNorm1d (helper.mlp_norm's default norm_type, helper.py:179-184)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        L, A, M, Hd = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim, cfg.hidden_dim
        self.prior_mlp = nn.Sequential(nn.Linear(Hd, M), nn.BatchNorm1d(M), nn.ELU(), nn.Linear(M, L))
        if not getattr(cfg, "norm_cell", True):
            raise NotImplementedError("tdmpc_b200 covers the NormGRUCell dynamics (cfg.norm_cell=True) only")
        self.gru_cell = NormGRUCell(L + A, Hd)

    def init_hidden_state(self, batch_size, device):
        return torch.zeros((batch_size, self.cfg.hidden_dim), device=device)

    def forward(self, obs_embed, action, h_prev):
        h = self.gru_cell(torch.cat([obs_embed, action], dim=-1), h_prev)
        return self.prior_mlp(h), h


class DSSMRnn(TOLD):
    """Model of the DRNN iCEM agent (src/algorithm/tdmpc_icem_similarity_drnn.py:14-77): GRU latent dynamics,
    reward read from the GRU state, `next(z, a, h_prev) -> (z', h', r)`.  Same state_dict keys as the reference."""

    def __init__(self, cfg):
        nn.Module.__init__(self)
        self.cfg = cfg
        L, A, M, E, D, Hd = cfg.latent_dim, cfg.action_dim, cfg.mlp_dim, cfg.enc_dim, cfg.obs_shape[0], cfg.hidden_dim
        if getattr(cfg, "modality", "state") != "state":
            raise NotImplementedError("tdmpc_b200 covers the state-observation encoder only")
        self._dynamics = DGruDyna(cfg)
        self._reward = _mlp(Hd, M, 1)
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
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
        for mod in (self._reward, self._Q1, self._Q2):
            mod[-1].weight.data.fill_(0)
            mod[-1].bias.data.fill_(0)

    def next(self, z, a, h_prev):
        z, hidden = self._dynamics(z, a, h_prev)
        return z, hidden, self._reward(hidden)

    def init_hidden_state(self, batch_size, device):
        return self._dynamics.init_hidden_state(batch_size, device)

    def pred_z(self, z):
        return self._predictor(z)

"""`ReplayBuffer` with the reference's surface (src/algorithm/helper.py:435-534) whose `sample()` runs on the
device through the C ABI (tdmpc_replay_sample / tdmpc_replay_gather, include/tdmpc_b200.h).

The reference moves the whole priority vector to the host on every update (`probs.cpu().numpy()`, helper.py:511)
and draws the batch with `np.random.choice`; here the power-law probabilities, the float64 cdf walk, numpy's
redraw rounds for `replace=False`, the importance weights and the horizon+1 step windows are CUDA kernels and
nothing but two status words crosses PCIe.  State modality only (pixels are out of scope).
There is no CPU fallback: constructing the buffer without a CUDA device raises.
"""
import ctypes as C
from copy import deepcopy

import torch

from . import _lib


class ReplayBuffer:
    """Drop-in for helper.ReplayBuffer: `buffer += episode`, `buffer.sample()`, `buffer.update_priorities(idxs, p)`.

    Differences a caller can see: `sample(uniforms=...)` optionally takes the float64 uniform stream
    (tests feed the one `np.random.seed(s)` would produce, which makes the indices those of the reference);
    without it a Philox stream keyed by (`seed`, number of sample() calls so far) is used."""

    MAX_ROUNDS = 32          # numpy loops until done; beyond this the draw is reported as failed
    ROUNDS_PER_CALL = 2      # redraw rounds enqueued before the status words are read

    def __init__(self, cfg, latent_plan=False, seed=0, device=None):
        if cfg.modality != 'state':
            raise NotImplementedError("tdmpc_b200.ReplayBuffer: state modality only")
        if not torch.cuda.is_available():
            raise RuntimeError("tdmpc_b200.ReplayBuffer needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.cfg = deepcopy(cfg)
        self.device = torch.device(device if device is not None else 'cuda')
        self.capacity = min(cfg.train_steps, cfg.max_buffer_size)                                   # helper.py:444
        obs_shape = tuple(cfg.obs_shape)
        f32 = dict(dtype=torch.float32, device=self.device)
        self._obs = torch.zeros((self.capacity + 1, *obs_shape), **f32)
        self._last_obs = torch.zeros((self.capacity // cfg.episode_length, *obs_shape), **f32)
        self._action = torch.zeros((self.capacity, cfg.action_dim), **f32)
        self._reward = torch.zeros((self.capacity,), **f32)
        self._priorities = torch.ones((self.capacity,), **f32)
        self._eps = 1e-6
        self._full = False
        self.idx = 0
        self.batch_size = cfg.batch_size
        self.horizon = cfg.horizon if latent_plan else cfg.env_horizon                              # helper.py:459-464
        if self.horizon > cfg.horizon:
            # only the last cfg.horizon transitions of an episode get priority 0 (helper.py:479-480): a longer window can
            # start close enough to the end of the buffer to run past it.  The reference then dies with an IndexError in
            # sample(); the gather kernel would read out of bounds silently, so refuse up front.
            raise ValueError(f"ReplayBuffer: sampling window {self.horizon} exceeds cfg.horizon {cfg.horizon}")
        self._seed, self._calls = int(seed), 0
        self._ws = torch.empty(int(self._lib.tdmpc_replay_workspace_bytes(self.capacity, self.batch_size)),
                               dtype=torch.uint8, device=self.device)
        self._status = torch.zeros(2, dtype=torch.int32, device=self.device)
        self._tail = torch.arange(cfg.episode_length, device=self.device) >= cfg.episode_length - cfg.horizon

    def __add__(self, episode):
        self.add(episode)
        return self

    def add(self, episode):
        """helper.py:470-486; the running maximum stays on the device (the reference reads it back with .item())."""
        ep = self.cfg.episode_length
        i = self.idx
        obs = torch.as_tensor(episode.obs, dtype=torch.float32).to(self.device)
        self._obs[i:i + ep] = obs[:-1]
        self._last_obs[i // ep] = obs[-1]
        self._action[i:i + ep] = torch.as_tensor(episode.action, dtype=torch.float32).to(self.device)
        self._reward[i:i + ep] = torch.as_tensor(episode.reward, dtype=torch.float32).to(self.device)
        if self._full:
            max_priority = self._priorities.max()
        elif i == 0:
            max_priority = torch.ones((), device=self.device)
        else:
            max_priority = self._priorities[:i].max()
        new = max_priority.expand(ep).clone()
        new[self._tail] = 0                       # the last cfg.horizon transitions of an episode are never drawn
        self._priorities[i:i + ep] = new
        self.idx = (i + ep) % self.capacity
        self._full = self._full or self.idx == 0

    def update_priorities(self, idxs, priorities):
        self._priorities[idxs] = priorities.squeeze(1).to(self.device) + self._eps                  # helper.py:488-489

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"tdmpc_b200 replay: {self._lib.tdmpc_last_error(None).decode()} (status {rc})")

    def _sample_args(self, idxs, weights, uniforms=None, priorities=None):
        """tdmpc_replay_sample_args for the buffer's current fill state (keeps `uniforms` alive on the struct)."""
        total = self.capacity if self._full else self.idx
        if total <= 0:
            raise RuntimeError("tdmpc_b200 replay: sample() on an empty buffer")
        a = _lib.ReplaySampleArgs()
        a.priorities = (self._priorities if priorities is None else priorities).data_ptr()
        a.total, a.batch = total, self.batch_size
        a.replace = int(not self._full)                                                             # helper.py:511
        a.per_alpha, a.per_beta = float(self.cfg.per_alpha), float(self.cfg.per_beta)
        if uniforms is not None:
            a._keep = uniforms = torch.as_tensor(uniforms, dtype=torch.float64).to(self.device).contiguous()
            a.uniforms, a.n_uniforms = uniforms.data_ptr(), uniforms.numel()
        a.seed, a.call = self._seed, self._calls
        a.max_rounds = self.ROUNDS_PER_CALL
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        a.idxs, a.weights, a.status = idxs.data_ptr(), weights.data_ptr(), self._status.data_ptr()
        return a

    def sample_indices(self, uniforms=None):
        """helper.py:507-513 on the device: (idxs int64 [batch], weights float32 [batch])."""
        idxs = torch.empty(self.batch_size, dtype=torch.int64, device=self.device)
        weights = torch.empty(self.batch_size, dtype=torch.float32, device=self.device)
        a = self._sample_args(idxs, weights, uniforms)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        found, rounds = 0, 0
        while True:
            a.resume = rounds
            self._check(self._lib.tdmpc_replay_sample(C.byref(a), stream))
            before = found
            found, now = self._status.tolist()         # the one host read (the reference syncs on the whole vector here)
            stuck = now == rounds or found == before   # out of injected uniforms / nothing drawable left
            rounds = now
            if found >= self.batch_size or stuck or rounds >= self.MAX_ROUNDS:
                break
        self._calls += 1
        if found < self.batch_size:
            raise RuntimeError(f"tdmpc_b200 replay: {found} of {self.batch_size} distinct indices after {rounds} rounds "
                               "(fewer drawable transitions than batch_size, or too few uniforms supplied)")
        self.last_rounds = rounds
        return idxs, weights

    def gather(self, idxs):
        """helper.py:515-529 on the device."""
        B, H = idxs.numel(), self.horizon
        D, A = self._obs.shape[1], self._action.shape[1]
        f32 = dict(dtype=torch.float32, device=self.device)
        obs = torch.empty((B, D), **f32)
        next_obs = torch.empty((H + 1, B, D), **f32)
        action = torch.empty((H + 1, B, A), **f32)
        reward = torch.empty((H + 1, B), **f32)
        g = _lib.ReplayGatherArgs()
        g.idxs, g.batch, g.horizon, g.obs_dim, g.action_dim = idxs.data_ptr(), B, H, D, A
        g.episode_length = self.cfg.episode_length
        g.obs_buf, g.last_obs_buf = self._obs.data_ptr(), self._last_obs.data_ptr()
        g.action_buf, g.reward_buf = self._action.data_ptr(), self._reward.data_ptr()
        g.obs, g.next_obs, g.action, g.reward = obs.data_ptr(), next_obs.data_ptr(), action.data_ptr(), reward.data_ptr()
        self._check(self._lib.tdmpc_replay_gather(C.byref(g), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return obs, next_obs, action, reward

    def sample(self, uniforms=None):
        """-> (obs [B,D], next_obs [H+1,B,D], action [H+1,B,A], reward [H+1,B,1], idxs [B], weights [B]), helper.py:506-534."""
        idxs, weights = self.sample_indices(uniforms)
        obs, next_obs, action, reward = self.gather(idxs)
        return obs, next_obs, action, reward.unsqueeze(2), idxs, weights

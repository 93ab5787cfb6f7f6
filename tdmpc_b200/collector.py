"""Batched environment-stepping loop around the CUDA planner.

The reference collects experience one environment at a time (src/train.py:88-99):

    obs = env.reset(); episode = Episode(cfg, obs)
    while not episode.done:
        action, plan_metrics = agent.plan(obs, step=step, t0=episode.first)
        obs, reward, done, _ = env.step(action.cpu().numpy())
        episode += (obs, action, reward, done)

i.e. per environment step one host->device copy of the observation, ~1500 aten dispatches, and one
device->host copy of the action.  `PlanLoop` is the same loop for B environments that advance in
lock-step: per step ONE pinned host->device copy of the [B, D] observations (+ the [B] t0 flags), ONE
replay of a CUDA graph that captured `tdmpc_plan` (all kernels of the plan, no host work in between),
and ONE device->host copy of the [B, A] actions and [B, 2] metrics into pinned memory.  Every replay
draws fresh noise: the graph captured a device-side call counter (`tdmpc_plan_args.call_dev`) that the
library increments at the end of each plan.  `t0` is per environment, so episodes may reset
independently (`_prev_mean` of that environment is zeroed, tdmpc.py:123-126).

`BatchedEpisodes` is the storage side: the `Episode` fields of the reference (algorithm/helper.py:383-415)
for B environments in pinned host memory, written once per step, handed out per environment in the
reference's layout so that `buffer += episode` keeps working.
"""
import numpy as np
import torch


def _pinned(*shape, dtype=torch.float32):
    """Page-locked host tensor (plain host memory where there is no CUDA driver: CPU-only unit tests of the storage)."""
    t = torch.empty(*shape, dtype=dtype)
    return t.pin_memory() if torch.cuda.is_available() else t


class PlanLoop:
    """B lock-step environments -> actions, one H2D + one graph replay + one D2H per step.

    planner: tdmpc_b200.Planner (weights already synced); plan_kwargs: the keyword arguments of
    Planner.plan that stay fixed (horizon, num_samples, num_pi_trajs, num_elites, iterations, discount,
    temperature, momentum, min_std, std_floor, ...)."""

    def __init__(self, planner, batch, plan_kwargs, eval_mode=False, use_graph=True, seed=None, call0=0):
        self.pl, self.B, self.kw = planner, int(batch), dict(plan_kwargs)
        self.eval_mode, self.use_graph = bool(eval_mode), bool(use_graph)
        dev = planner.device
        D, A, H = planner.D, planner.A, int(self.kw["horizon"])
        self.seed = planner.seed if seed is None else int(seed)
        self.call0 = int(call0)
        # pinned host side of the ring (one slot: the step is synchronous at its end)
        self.obs_h = _pinned(self.B, D)
        self.t0_h = _pinned(self.B, dtype=torch.uint8).fill_(1)
        self.act_h = _pinned(self.B, A)
        self.met_h = _pinned(self.B, 2)
        # device side
        self.obs_d = torch.zeros(self.B, D, device=dev)
        self.t0_d = torch.ones(self.B, dtype=torch.uint8, device=dev)
        self.prev_mean = torch.zeros(self.B, H, A, device=dev)
        self.out = {"action": torch.zeros(self.B, A, device=dev), "metrics": torch.zeros(self.B, 2, device=dev)}
        self.call_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = None
        self.steps = 0
        self.h2d_bytes_per_step = self.obs_h.numel() * 4 + self.t0_h.numel()
        self.d2h_bytes_per_step = self.act_h.numel() * 4 + self.met_h.numel() * 4

    # everything one step enqueues; captured once, replayed afterwards
    def _enqueue(self):
        self.obs_d.copy_(self.obs_h, non_blocking=True)
        self.t0_d.copy_(self.t0_h, non_blocking=True)
        self.pl.plan(self.obs_d, eval_mode=self.eval_mode, t0=self.t0_d, prev_mean=self.prev_mean, has_prev=True,
                     out=self.out, seed=self.seed, call=self.call0, call_dev=self.call_dev, **self.kw)
        self.act_h.copy_(self.out["action"], non_blocking=True)
        self.met_h.copy_(self.out["metrics"], non_blocking=True)

    def _capture(self):
        # workspace reservation and weight packing must happen outside the capture: one plain call first (its effect on
        # the planner state is undone: prev_mean and the call counter are reset)
        keep_prev = self.prev_mean.clone()
        with torch.cuda.stream(self.stream):
            self._enqueue()
        self.stream.synchronize()
        self.prev_mean.copy_(keep_prev)
        self.call_dev.zero_()
        torch.cuda.synchronize(self.pl.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            self._enqueue()
        self.graph = g

    def step(self, obs, t0):
        """obs [B, D] array-like (host), t0 [B] bool/uint8 or a scalar -> (actions [B, A], metrics [B, 2]) as numpy views
        of the pinned result buffers (valid until the next step)."""
        self.obs_h.copy_(torch.as_tensor(np.asarray(obs, dtype=np.float32)).reshape(self.B, -1))
        t0a = np.asarray(t0)
        self.t0_h.copy_(torch.as_tensor(np.broadcast_to(t0a.astype(np.uint8), (self.B,)).copy()))
        if self.use_graph:
            if self.graph is None:
                self._capture()
            with torch.cuda.stream(self.stream):
                self.graph.replay()
        else:
            with torch.cuda.stream(self.stream):
                self._enqueue()
        self.stream.synchronize()
        self.steps += 1
        return self.act_h.numpy(), self.met_h.numpy()


class BatchedEpisodes:
    """The reference's `Episode` (algorithm/helper.py:383-415) for B lock-step environments, in pinned host memory.

    add(obs, action, reward, done) appends one transition per environment; episode(i) returns an object with the fields
    the reference's ReplayBuffer reads (`obs` [T+1, D], `action` [T, A], `reward` [T], `cumulative_reward`, `done`,
    `first`, len())."""

    def __init__(self, batch, episode_length, obs_dim, action_dim, init_obs):
        self.B, self.T = int(batch), int(episode_length)
        self.obs = _pinned(self.B, self.T + 1, obs_dim)
        self.action = _pinned(self.B, self.T, action_dim)
        self.reward = _pinned(self.B, self.T)
        self.cumulative_reward = np.zeros(self.B, dtype=np.float64)
        self.done = np.zeros(self.B, dtype=bool)
        self._idx = 0
        self.obs[:, 0] = torch.as_tensor(np.asarray(init_obs, dtype=np.float32)).reshape(self.B, obs_dim)

    def __len__(self):
        return self._idx

    @property
    def first(self):
        return self._idx == 0

    def add(self, obs, action, reward, done):
        if self._idx >= self.T:
            raise IndexError("episode is full")
        self.obs[:, self._idx + 1] = torch.as_tensor(np.asarray(obs, dtype=np.float32)).reshape(self.B, -1)
        self.action[:, self._idx] = torch.as_tensor(np.asarray(action, dtype=np.float32)).reshape(self.B, -1)
        r = np.asarray(reward, dtype=np.float32).reshape(self.B)
        self.reward[:, self._idx] = torch.as_tensor(r)
        self.cumulative_reward += r
        self.done = np.broadcast_to(np.asarray(done, dtype=bool), (self.B,)).copy()
        self._idx += 1
        return self

    __iadd__ = lambda self, tr: self.add(*tr)

    def episode(self, i):
        return EpisodeView(self.obs[i, :self._idx + 1], self.action[i, :self._idx], self.reward[i, :self._idx],
                           float(self.cumulative_reward[i]), bool(self.done[i]))


class EpisodeView:
    """One environment's episode in the reference's `Episode` layout (helper.py:383-415): what `ReplayBuffer.__add__` reads."""

    def __init__(self, obs, action, reward, cumulative_reward, done):
        self.obs, self.action, self.reward = obs, action, reward
        self.cumulative_reward, self.done = cumulative_reward, done
        self._idx = action.shape[0]

    def __len__(self):
        return self._idx

    @property
    def first(self):
        return self._idx == 0

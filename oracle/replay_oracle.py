"""CPU oracle of the prioritised replay sampler -- TEST INFRASTRUCTURE, never on the product path.

Restates `ReplayBuffer.sample` of the reference (src/algorithm/helper.py:506-534) with torch CPU ops,
and the one third-party routine it calls, `numpy.random.choice` (legacy `RandomState.choice`,
numpy/random/mtrand.pyx), with the uniform draws INJECTED so that the CUDA sampler can be fed the
same stream.  Pinned by tests/golden/replay_*.npz: outputs of the reference class itself, run on
CPU by tests/golden/make_golden.py, and by `numpy.random.RandomState.choice` directly
(tests/test_replay_oracle.py).
"""
import numpy as np
import torch


def choice(total, size, p, replace, uniforms):
    """np.random.choice(total, size, p=p, replace=replace) with `uniforms` standing in for the
    generator's random_sample() stream.  Returns (indices int64 [size], number of uniforms consumed).

    mtrand.pyx: cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(side='right').  Without replacement numpy
    loops: draw `size - n_uniq` values, zero the probabilities of what was found, keep the first
    occurrence of each new index in draw order, until `size` distinct indices exist."""
    p = np.asarray(p).astype(np.float64)            # numpy casts the float32 probabilities to double
    uniforms = np.asarray(uniforms, dtype=np.float64)
    if replace:
        cdf = p.cumsum()
        cdf /= cdf[-1]
        return cdf.searchsorted(uniforms[:size], side="right").astype(np.int64), size
    if size > np.count_nonzero(p > 0):
        raise ValueError("Fewer non-zero entries in p than size")
    n_uniq, used = 0, 0
    p = p.copy()
    found = np.zeros(size, dtype=np.int64)
    while n_uniq < size:
        m = size - n_uniq
        x = uniforms[used:used + m]
        if len(x) < m:
            raise ValueError("ran out of injected uniforms")
        used += m
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, unique_indices = np.unique(new, return_index=True)
        unique_indices.sort()
        new = new.take(unique_indices)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found, used


def sample(state, uniforms):
    """ReplayBuffer.sample (helper.py:506-534), state modality.

    state: dict with the reference's buffers as CPU tensors -- obs [cap+1, D], last_obs [cap/ep, D],
    action [cap, A], reward [cap], priorities [cap] -- and the scalars idx, full, batch_size, horizon,
    episode_length, per_alpha, per_beta.  Returns the reference's tuple plus the number of uniforms used."""
    pri = state["priorities"] if state["full"] else state["priorities"][:state["idx"]]
    probs = pri ** state["per_alpha"]                                            # helper.py:507
    probs /= probs.sum()                                                         # :508
    total = len(probs)                                                           # :509
    idx_np, used = choice(total, state["batch_size"], probs.numpy(), not state["full"], uniforms)   # :510-511
    idxs = torch.from_numpy(idx_np)
    weights = (total * probs[idxs]) ** (-state["per_beta"])                      # :512
    weights /= weights.max()                                                     # :513
    H, B, ep = state["horizon"], state["batch_size"], state["episode_length"]
    obs = state["obs"][idxs]                                                     # :515
    next_obs = torch.empty((H + 1, B, state["obs"].shape[1]), dtype=torch.float32)
    action = torch.empty((H + 1, B, state["action"].shape[1]), dtype=torch.float32)
    reward = torch.empty((H + 1, B), dtype=torch.float32)
    for t in range(H + 1):                                                       # :522-526
        _idxs = idxs + t
        next_obs[t] = state["obs"][_idxs + 1]
        action[t] = state["action"][_idxs]
        reward[t] = state["reward"][_idxs]
    mask = (_idxs + 1) % ep == 0                                                 # :528 (uses the last step's _idxs)
    next_obs[-1, mask] = state["last_obs"][_idxs[mask] // ep]                    # :529
    return obs, next_obs, action, reward.unsqueeze(2), idxs, weights, used


class ReplayState(dict):
    """The storage half of the reference buffer (helper.py:441-489) as CPU tensors, so that `sample` above
    can be driven without the reference: __init__ (:441-464), add (:470-486), update_priorities (:488-489)."""

    def __init__(self, cfg, latent_plan=False):
        cap = min(cfg.train_steps, cfg.max_buffer_size)                                       # :444
        D, A = cfg.obs_shape[0], cfg.action_dim
        super().__init__(
            obs=torch.zeros((cap + 1, D)), last_obs=torch.zeros((cap // cfg.episode_length, D)),
            action=torch.zeros((cap, A)), reward=torch.zeros((cap,)), priorities=torch.ones((cap,)),
            idx=0, full=False, capacity=cap, batch_size=cfg.batch_size,
            horizon=cfg.horizon if latent_plan else cfg.env_horizon,                          # :459-464
            episode_length=cfg.episode_length, per_alpha=cfg.per_alpha, per_beta=cfg.per_beta, plan_horizon=cfg.horizon)

    def add(self, obs, action, reward):
        """obs [ep+1, D], action [ep, A], reward [ep] of one episode."""
        s, ep = self, self["episode_length"]
        i = s["idx"]
        s["obs"][i:i + ep] = obs[:-1]
        s["last_obs"][i // ep] = obs[-1]
        s["action"][i:i + ep] = action
        s["reward"][i:i + ep] = reward
        if s["full"]:
            max_priority = s["priorities"].max().item()
        else:
            max_priority = 1. if i == 0 else s["priorities"][:i].max().item()
        new = torch.full((ep,), max_priority)
        new[torch.arange(ep) >= ep - s["plan_horizon"]] = 0          # the last cfg.horizon transitions can never be drawn
        s["priorities"][i:i + ep] = new
        s["idx"] = (i + ep) % s["capacity"]
        s["full"] = s["full"] or s["idx"] == 0

    def update_priorities(self, idxs, priorities):
        self["priorities"][idxs] = priorities.squeeze(1) + 1e-6                               # :489 (_eps)

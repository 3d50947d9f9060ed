"""Device-resident synthetic environment feed for benchmarking the engine without env processes.

The reference's rollout talks to one OS process per environment over pipes (worker.py); that stays
the default transport here (``worker.Worker``).  This feed is the "inputs already resident in HBM"
variant used by ``bench.py`` for its device-timed ``value``: observations for every (step, worker)
are pre-generated ON the GPU, rewards / episode ends are a pre-computed host schedule with the same
distribution as ``SyntheticEnv`` (uniform obs in [0,1), N(0, 0.1) rewards, episode lengths uniform in
[min, max]), so the rollout loop needs no host<->device traffic and no per-step synchronisation.
"""
import numpy as np
import torch


class SyntheticDeviceFeed:
    def __init__(self, n_workers, worker_steps, obs_shape, max_episode_steps, min_episode_steps=None, seed=0, device="cuda"):
        self.W, self.T = n_workers, worker_steps
        self.max_len = int(max_episode_steps)
        self.min_len = int(min_episode_steps) if min_episode_steps else max(1, self.max_len // 4)
        self.rng = np.random.default_rng(seed)
        g = torch.Generator(device=device).manual_seed(seed)
        # T+1 observation slabs: slab t is what the workers see at step t; slab T seeds the next update
        self.obs_all = torch.rand((worker_steps + 1, n_workers) + tuple(obs_shape), generator=g, device=device)
        self.remaining = self.rng.integers(self.min_len, self.max_len + 1, size=n_workers)
        self.ep_len = self.remaining.copy()
        self.ep_ret = np.zeros(n_workers, dtype=np.float64)
        self.rewards = np.zeros((worker_steps, n_workers), dtype=np.float32)
        self.dones = np.zeros((worker_steps, n_workers), dtype=bool)
        self.infos = [[] for _ in range(worker_steps)]

    def begin_update(self):
        """Draw this update's reward / done schedule (episodes continue across updates)."""
        T, W = self.T, self.W
        self.rewards[:] = self.rng.normal(0.0, 0.1, size=(T, W)).astype(np.float32)
        self.dones[:] = False
        self.infos = [[] for _ in range(T)]
        for w in range(W):
            t = 0
            while True:
                end = t + int(self.remaining[w]) - 1          # step index at which this episode ends
                if end >= T:
                    self.ep_ret[w] += float(self.rewards[t:, w].sum())
                    self.remaining[w] -= (T - t)
                    break
                self.dones[end, w] = True
                self.ep_ret[w] += float(self.rewards[t:end + 1, w].sum())
                self.infos[end].append((w, {"reward": self.ep_ret[w], "length": int(self.ep_len[w])}))
                self.ep_ret[w] = 0.0
                self.ep_len[w] = self.remaining[w] = int(self.rng.integers(self.min_len, self.max_len + 1))
                t = end + 1
                if t >= T:
                    break

    def obs(self, t):
        return self.obs_all[t]

    def step(self, t):
        """(rewards (W,), dones (W,), [(worker, info), ...]) for step t."""
        return self.rewards[t], self.dones[t], self.infos[t]

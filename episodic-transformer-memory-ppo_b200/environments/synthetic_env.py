"""Synthetic benchmark environment: observations of a chosen shape drawn uniformly
from [0, 1), Gaussian rewards, random episode lengths.  It has no task to learn -- it exists so the
"Minigrid 3x84x84, memory_length=128" configurations of BASELINE.json can run at all: the real
Minigrid-Memory environment caps episodes at 96 steps (reference environments/minigrid_env.py:17),
below ``memory_length``, which the reference's window table cannot represent (trainer.py:89).

The interface is the reference's gym-style protocol (environments/minigrid_env.py:38-85):
``observation_space.shape``, ``action_space.n``, ``max_episode_steps``, ``reset() -> obs``,
``step(action) -> (obs, reward, done, info)`` with ``info = {"reward", "length"}`` on the last step of
an episode and ``None`` otherwise.  No gym dependency.
"""
import numpy as np


class _Box:
    def __init__(self, shape, low=0.0, high=1.0, dtype=np.float32):
        self.shape = tuple(shape)
        self.low, self.high, self.dtype = low, high, dtype


class _Discrete:
    def __init__(self, n):
        self.n = int(n)


class SyntheticEnv:
    def __init__(self, obs_shape=(3, 84, 84), n_actions=3, max_episode_steps=256, min_episode_steps=None, seed=0):
        self._obs_shape = tuple(obs_shape)
        self._n_actions = int(n_actions)
        self._max_steps = int(max_episode_steps)
        self._min_steps = int(min_episode_steps) if min_episode_steps else max(1, self._max_steps // 4)
        assert 1 <= self._min_steps <= self._max_steps
        self._rng = np.random.default_rng(seed)
        self._t = 0
        self._len = self._max_steps
        self._ret = 0.0
        self._out = None

    @property
    def observation_space(self):
        return _Box(self._obs_shape)

    @property
    def action_space(self):
        return _Discrete(self._n_actions)

    @property
    def max_episode_steps(self):
        return self._max_steps

    def set_observation_buffer(self, out):
        """Optional: draw every observation directly into ``out`` (a float32 array of the observation shape, e.g. this
        worker's slot of the trainer's shared pinned slab) instead of allocating a new array per step; same random stream."""
        assert out.shape == self._obs_shape and out.dtype == np.float32
        self._out = out

    def _obs(self):
        if self._out is not None:
            return self._rng.random(self._obs_shape, dtype=np.float32, out=self._out)
        return self._rng.random(self._obs_shape, dtype=np.float32)

    def reset(self, **kwargs):
        self._t = 0
        self._ret = 0.0
        self._len = int(self._rng.integers(self._min_steps, self._max_steps + 1))
        return self._obs()

    def step(self, action):
        self._t += 1
        reward = float(np.float32(self._rng.normal(0.0, 0.1)))
        self._ret += reward
        done = self._t >= self._len
        info = {"reward": self._ret, "length": self._t} if done else None
        return self._obs(), reward, done, info

    def render(self):
        return None

    def close(self):
        return None

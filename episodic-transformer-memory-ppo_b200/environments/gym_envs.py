"""Adapters for the reference's gym-backed environments (reference environments/cartpole_env.py,
minigrid_env.py, memory_gym_env.py) on the engine's gym-style protocol: ``observation_space.shape``,
``action_space.n``, ``max_episode_steps``, ``reset() -> obs``, ``step(action) -> (obs, reward, done,
info)`` with ``info = {"reward", "length"}`` on the last step of an episode and ``None`` otherwise.

The third-party packages (gym, gym_minigrid, gymnasium + memory_gym) are imported when an adapter is
constructed, never at module import, so the training engine does not depend on them; a missing
package raises ``MissingEnvDependency`` naming the package to install.  None of them is installed in
the build image, so these adapters are exercised with stand-in packages in tests/test_host.py."""
import importlib
import time

import numpy as np


class MissingEnvDependency(ImportError):
    pass


def _require(package, env_type):
    try:
        return importlib.import_module(package)
    except ImportError as e:
        raise MissingEnvDependency(
            "environment type %r needs the third-party package %r (pip name: %s), which is not installed; the B200 engine "
            "itself only needs the gym-style protocol, e.g. the built-in `Synthetic` and `PocMemoryEnv` types"
            % (env_type, package, {"gym_minigrid": "gym-minigrid==1.0.2", "gym": "gym==0.18.3", "memory_gym": "memory-gym",
                                   "gymnasium": "gymnasium"}.get(package, package))) from e


def _chw(image):
    """(H, W, C) image -> (C, W', H') as the reference does with two swapaxes (0,2 then 2,1), i.e. channels first."""
    return np.swapaxes(np.swapaxes(image, 0, 2), 2, 1)


class _EpisodeStats:
    """Accumulates the per-episode reward list behind the ``info = {"reward", "length"}`` convention."""

    def __init__(self):
        self.rewards = []

    def begin(self):
        self.rewards = []

    def add(self, reward, done):
        self.rewards.append(reward)
        return {"reward": sum(self.rewards), "length": len(self.rewards)} if done else None


class CartPole:
    """CartPole-v0 with optional velocity masking (partial observability) and rewards scaled by 1/100
    (reference environments/cartpole_env.py:5-43)."""

    def __init__(self, mask_velocity=False):
        gym = _require("gym", "CartPoleMasked" if mask_velocity else "CartPole")
        self._env = gym.make("CartPole-v0")
        self.max_episode_steps = self._env.spec.max_episode_steps
        self._keep = np.array([1, 0, 1, 0] if mask_velocity else [1, 1, 1, 1], dtype=np.float32)
        self._stats = _EpisodeStats()

    observation_space = property(lambda self: self._env.observation_space)
    action_space = property(lambda self: self._env.action_space)

    def reset(self):
        self._stats.begin()
        return self._env.reset() * self._keep

    def step(self, action):
        obs, reward, done, _ = self._env.step(action[0])
        return obs * self._keep, reward / 100.0, done, self._stats.add(reward, done)

    def render(self):
        self._env.render()
        time.sleep(0.033)

    def close(self):
        self._env.close()


class Minigrid:
    """gym-minigrid with a reduced, RGB partial view, float CHW observations in [0, 1] and a hard step limit
    (reference environments/minigrid_env.py:8-85).  Memory tasks: 3x3 view at 28 px tiles (84x84), 3 actions, 96 steps;
    other tasks: 7x7 view at 8 px tiles (56x56), the env's own actions, 64 steps."""

    def __init__(self, name):
        gym = _require("gym", "Minigrid")
        wrappers = _require("gym_minigrid.wrappers", "Minigrid")
        memory_task = "Memory" in name
        view, self.tile_size, self.max_episode_steps = (3, 28, 96) if memory_task else (7, 8, 64)
        env = gym.make(name)
        self._action_space = gym.spaces.Discrete(3) if memory_task else env.action_space
        env = wrappers.RGBImgPartialObsWrapper(wrappers.ViewSizeWrapper(env, view), tile_size=self.tile_size)
        self._env = wrappers.ImgObsWrapper(env)
        self._observation_space = gym.spaces.Box(low=0, high=1.0, shape=(3, view * self.tile_size, view * self.tile_size),
                                                 dtype=np.float32)
        self._stats, self.t = _EpisodeStats(), 0

    observation_space = property(lambda self: self._observation_space)
    action_space = property(lambda self: self._action_space)

    def reset(self):
        self._env.seed(np.random.randint(0, 999))
        self.t = 0
        self._stats.begin()
        return _chw(self._env.reset().astype(np.float32) / 255.)

    def step(self, action):
        obs, reward, done, _ = self._env.step(action[0])
        done = done or self.t == self.max_episode_steps - 1
        self.t += 1
        return _chw(obs.astype(np.float32) / 255.), reward, done, self._stats.add(reward, done)

    def render(self):
        img = self._env.render(tile_size=96)
        time.sleep(0.5)
        return img

    def close(self):
        self._env.close()


class MemoryGymWrapper:
    """memory-gym (gymnasium API) environments -- SearingSpotlights, MortarMayhem(-Grid), MysteryPath(-Grid) -- with CHW
    observations scaled to [0, 1]; episode ``info`` comes from the environment itself (reference
    environments/memory_gym_env.py:10-119)."""

    def __init__(self, env_name, reset_params=None, realtime_mode=False):
        gym = _require("gymnasium", env_name)
        _require("memory_gym", env_name)
        self._reset_params = {"start-seed": 0, "num-seeds": 100} if reset_params is None else reset_params
        self._env = gym.make(env_name, disable_env_checker=True, render_mode="human" if realtime_mode else None)
        h, w, c = self._env.observation_space.shape
        self._observation_space = gym.spaces.Box(low=0, high=1.0, shape=(c, w, h), dtype=np.float32)

    observation_space = property(lambda self: self._observation_space)
    action_space = property(lambda self: self._env.action_space)

    @property
    def max_episode_steps(self):
        self._env.reset()
        return self._env.max_episode_steps

    def reset(self, reset_params=None):
        params = self._reset_params if reset_params is None else reset_params
        seed = int(np.random.randint(params["start-seed"], params["start-seed"] + params["num-seeds"]))
        options = {k: v for k, v in params.items() if k not in ("start-seed", "num-seeds", "seed")}
        obs, _ = self._env.reset(seed=seed, options=options)
        return _chw(obs) / 255.0

    def step(self, action):
        if isinstance(action, (list, tuple, np.ndarray)) and len(action) == 1:
            action = action[0]
        obs, reward, done, _truncated, info = self._env.step(action)
        return _chw(obs) / 255.0, reward, done, info

    def render(self):
        self._env.render()

    def close(self):
        self._env.close()

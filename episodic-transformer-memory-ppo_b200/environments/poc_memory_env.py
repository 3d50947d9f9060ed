"""Proof-of-concept memory task on a 1-D corridor (same task and reward structure as the reference's
``environments/poc_memory_env.py``; written from its description, not copied).

The agent starts near the middle of a corridor on [-1, 1].  One end holds a rewarding goal, the other
a punishing one; which is which is visible in the observation ``[left_goal, position, right_goal]`` only
during the first two steps, while (with ``freeze=True``) the agent cannot move.  Afterwards the goal
entries read 0 and the agent has to walk to the end it remembers.  Reaching an end terminates the
episode with +/-(1 + min_steps * time_penalty); every other step costs ``time_penalty``.
"""
import numpy as np


class _Box:
    def __init__(self, shape):
        self.shape, self.low, self.high, self.dtype = tuple(shape), 0.0, 1.0, np.float32


class _Discrete:
    def __init__(self, n):
        self.n = n


class PocMemoryEnv:
    def __init__(self, step_size=0.2, glob=False, freeze=False, max_episode_steps=-1):
        self.freeze = bool(freeze)
        self._step = float(step_size)
        self.max_episode_steps = int(max_episode_steps)
        self._min_steps = int(1.0 / self._step) + 1
        self._time_penalty = 0.1
        self._show_steps = 2
        ticks = int(0.4 / self._step)
        if glob:
            lo, hi = -1 + self._step, 1
        else:
            lo = min(-2.0 * self._step, -ticks * self._step)
            hi = max(3.0 * self._step, self._step, (ticks + 1) * self._step)
        starts = np.arange(lo, hi, self._step).clip(-1 + self._step, 1 - self._step)
        self.possible_positions = [round(float(p), 2) for p in starts]
        self._pos, self._goals, self._t, self._rewards = 0.0, np.array([-1.0, 1.0]), 0, []

    @property
    def observation_space(self):
        return _Box((3,))

    @property
    def action_space(self):
        return _Discrete(2)

    def _observe(self, show):
        left, right = (self._goals[0], self._goals[1]) if show else (0.0, 0.0)
        return np.asarray([left, self._pos, right], dtype=np.float32)

    def reset(self, **kwargs):
        self._pos = float(np.random.choice(self.possible_positions))
        self._goals = np.asarray([-1.0, 1.0])[np.random.permutation(2)]
        self._t, self._rewards = 0, []
        return self._observe(True)

    def step(self, action):
        move = self._step if int(action[0]) == 1 else -self._step
        done = self.max_episode_steps > 0 and self._t >= self.max_episode_steps - 1
        showing = self._t < self._show_steps
        if showing and self.freeze:                      # goals visible, agent held in place
            self._t += 1
            self._rewards.append(0.0)
            return self._observe(True), 0.0, done, None
        self._pos = float(np.round(self._pos + move, 2))
        obs = self._observe(showing)
        reward, success = 0.0, False
        terminal_bonus = 1.0 + self._min_steps * self._time_penalty
        if self._pos == -1.0 or self._pos == 1.0:
            good = self._goals[0 if self._pos == -1.0 else 1] == 1.0
            reward += terminal_bonus if good else -terminal_bonus
            success, done = bool(good), True
        else:
            reward -= self._time_penalty
        self._rewards.append(reward)
        info = {"success": success, "reward": float(sum(self._rewards)), "length": len(self._rewards)} if done else None
        self._t += 1
        return obs, reward, done, info

    def render(self):
        cells = int(round(2.0 / self._step)) + 1
        at = int(round((self._pos + 1.0) / self._step))
        row = ["."] * cells
        row[0] = "+" if self._goals[0] > 0 else "-"
        row[-1] = "+" if self._goals[1] > 0 else "-"
        row[max(0, min(cells - 1, at))] = "a"
        print("".join(row), "(goals shown)" if self._t < self._show_steps else "")

    def close(self):
        return None

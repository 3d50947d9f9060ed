"""Helpers with the reference's names (utils.py): ``create_env``, ``polynomial_decay``,
``batched_index_select``, ``process_episode_info``, ``Module``."""
import numpy as np
import torch
from torch import nn

import trxl_native as native

def create_env(config, render=False):
    """Instantiate an environment from ``config["type"]`` (reference utils.py:11-30).  ``Synthetic``
    is this repo's benchmark environment; the gym-backed types import their
    dependency lazily so the training engine does not depend on gym."""
    kind = config["type"]
    if kind == "Synthetic":
        from environments.synthetic_env import SyntheticEnv
        return SyntheticEnv(obs_shape=tuple(config.get("obs_shape", (3, 84, 84))), n_actions=config.get("n_actions", 3),
                            max_episode_steps=config.get("max_episode_steps", 256),
                            min_episode_steps=config.get("min_episode_steps"), seed=config.get("seed", 0))
    if kind == "PocMemoryEnv":
        from environments.poc_memory_env import PocMemoryEnv
        return PocMemoryEnv(glob=False, freeze=True, max_episode_steps=32)
    # gym-backed types: the adapters import gym / gym_minigrid / gymnasium + memory_gym when constructed and raise
    # environments.gym_envs.MissingEnvDependency (an ImportError naming the package) when it is not installed
    if kind in ("CartPole", "CartPoleMasked"):
        from environments.gym_envs import CartPole
        return CartPole(mask_velocity=(kind == "CartPoleMasked"))
    if kind == "Minigrid":
        from environments.gym_envs import Minigrid
        return Minigrid(config["name"])
    if kind in ("SearingSpotlights", "MortarMayhem", "MortarMayhem-Grid", "MysteryPath", "MysteryPath-Grid"):
        from environments.gym_envs import MemoryGymWrapper
        return MemoryGymWrapper(env_name=config["name"], reset_params=config.get("reset_params"), realtime_mode=render)
    raise ValueError("unknown environment type %r" % kind)


def polynomial_decay(initial, final, max_decay_steps, power, current_step):
    """Polynomial schedule between ``initial`` and ``final`` (reference utils.py:32-50)."""
    if current_step > max_decay_steps or initial == final:
        return final
    frac = 1 - current_step / max_decay_steps
    return (initial - final) * (frac ** power) + final


def batched_index_select(input, dim, index):
    """``out[b, l, ...] = input[b, index[b, l], ...]`` (reference utils.py:52-75).  ``dim == 1`` on a
    CUDA fp32 tensor runs the native gather; the training engine itself never calls this (windows
    are read in place), it exists for API compatibility."""
    if dim == 1 and input.is_cuda and input.dtype == torch.float32 and input.dim() >= 2:
        src = input.contiguous()
        idx = index.to(src.device, torch.int64).contiguous()
        out = torch.empty((src.shape[0], idx.shape[1]) + tuple(src.shape[2:]), dtype=src.dtype, device=src.device)
        native.gather_window(src, idx, out)
        return out
    view = [1] * input.dim()
    view[0], view[dim] = index.shape[0], index.shape[1]
    expand = list(input.shape)
    expand[0], expand[dim] = -1, -1
    return torch.gather(input, dim, index.reshape(view).expand(expand))


def process_episode_info(episode_info):
    """Mean/std of every key of the finished-episode dicts; ``success`` also yields ``success_percent``
    (reference utils.py:77-95)."""
    result = {}
    if len(episode_info) == 0:
        return result
    for key in episode_info[0].keys():
        vals = [info[key] for info in episode_info]
        if key == "success":
            result[key + "_percent"] = np.sum(vals) / len(vals)
        result[key + "_mean"] = np.mean(vals)
        result[key + "_std"] = np.std(vals)
    return result


class Module(nn.Module):
    """nn.Module with gradient norm/mean helpers (reference utils.py:97-122)."""

    def _flat_grads(self):
        return [p.grad.view(-1) for _, p in self.named_parameters()]

    def grad_norm(self):
        g = self._flat_grads()
        return torch.linalg.norm(torch.cat(g)).item() if g else None

    def grad_mean(self):
        g = self._flat_grads()
        return torch.mean(torch.cat(g)).item() if g else None

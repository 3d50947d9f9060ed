"""Run one episode with a trained model (flag-compatible with the reference's enjoy.py: ``--model``).
Loads the reference's checkpoint format, a pickled ``(state_dict, config)`` tuple, and steps the policy
with the same per-step memory bookkeeping (reference enjoy.py:9-25,60-84) on the GPU engine."""
import argparse
import os
import pickle
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from model import ActorCriticModel  # noqa: E402
from trainer import build_mask_table, build_window_index_table  # noqa: E402
from utils import create_env  # noqa: E402


def init_transformer_memory(trxl_conf, max_episode_steps, device):
    """Initial episodic memory (1, M, B, D), the (L, L) mask table and the (M, L) window-index table."""
    memory = torch.zeros((1, max_episode_steps, trxl_conf["num_blocks"], trxl_conf["embed_dim"]), device=device)
    return memory, build_mask_table(trxl_conf["memory_length"]).to(device), \
        build_window_index_table(max_episode_steps, trxl_conf["memory_length"]).to(device)


def load_model(path, env, device):
    """Load the reference's checkpoint format -- a pickled ``(state_dict, config)`` tuple (reference enjoy.py:47-57,
    written by trainer._save_model) -- into an ActorCriticModel on ``device``.  Returns (model, config)."""
    with open(path, "rb") as f:
        state_dict, config = pickle.load(f)
    model = ActorCriticModel(config, env.observation_space, (env.action_space.n,), env.max_episode_steps)
    model.load_state_dict(state_dict)
    model.to(device).eval()
    return model, config


def run_episode(model, env, config=None, device=None, render=False, record=False):
    """Step one episode (reference enjoy.py:60-84).  ``model`` is an ActorCriticModel or the path of a saved
    ``(state_dict, config)`` pickle.  Returns ``(info, rewards)``; with ``record=True`` a dict that also holds every
    step's observation, value and normalised logits (used by the parity tests)."""
    device = torch.device("cuda") if device is None else torch.device(device)
    if isinstance(model, (str, os.PathLike)):
        model, config = load_model(model, env, device)
    trxl = config["transformer"]
    memory, mask_table, index_table = init_transformer_memory(trxl, env.max_episode_steps, device)
    L, t, done, info, rewards = trxl["memory_length"], 0, False, None, []
    trace = {"obs": [], "values": [], "logits": [], "actions": []}
    obs = env.reset()
    with torch.no_grad():
        while not done:
            obs_t = torch.tensor(np.expand_dims(obs, 0), dtype=torch.float32, device=device)
            idx = index_table[t].unsqueeze(0)
            window = memory[0, index_table[t]].unsqueeze(0)
            mask = mask_table[max(0, min(t, L - 1))].unsqueeze(0)
            if render:
                env.render()
            policy, value, new_memory = model(obs_t, window, mask, idx)
            memory[:, t] = new_memory
            action = [int(branch.sample().item()) for branch in policy]
            if record:
                trace["obs"].append(np.asarray(obs, dtype=np.float32).copy())
                trace["values"].append(float(value[0]))
                trace["logits"].append(policy[0].logits[0].cpu().numpy())
                trace["actions"].append(action)
            obs, reward, done, info = env.step(action)
            rewards.append(reward)
            t += 1
    if record:
        trace.update(info=info, rewards=rewards, length=t)
        return trace
    return info, rewards


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="./models/run.nn", help="path to the trained model ((state_dict, config) pickle)")
    args = ap.parse_args(argv)
    device = torch.device("cuda")          # the engine has no CPU path
    with open(args.model, "rb") as f:
        _, config = pickle.load(f)
    env = create_env(config["environment"], render=True)
    model, config = load_model(args.model, env, device)
    info, _ = run_episode(model, env, config, device, render=True)
    print("Episode length: " + str(info["length"]))
    print("Episode reward: " + str(info["reward"]))
    env.close()


if __name__ == "__main__":
    main()

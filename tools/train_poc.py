#!/usr/bin/env python
"""End-to-end learning check on the proof-of-concept memory task (the reference's default config):
trains PPO+TrXL with this engine and prints the success rate per block of updates."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "episodic-transformer-memory-ppo_b200")
sys.path.insert(0, PKG)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--updates", type=int, default=120)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    from trainer import PPOTrainer
    from utils import polynomial_decay, process_episode_info
    from yaml_parser import YamlParser
    cfg = YamlParser(os.path.join(PKG, "configs", "poc_memory_env.yaml")).get_config()
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    os.chdir("/tmp")
    tr = PPOTrainer(cfg, run_id="poc", device=torch.device("cuda:0"), summary_writer=False)
    curve = []
    for update in range(args.updates):
        lr = polynomial_decay(**{k: cfg["learning_rate_schedule"][k] for k in ("initial", "final", "max_decay_steps", "power")}, current_step=update)
        beta = polynomial_decay(**{k: cfg["beta_schedule"][k] for k in ("initial", "final", "max_decay_steps", "power")}, current_step=update)
        clip = polynomial_decay(**{k: cfg["clip_range_schedule"][k] for k in ("initial", "final", "max_decay_steps", "power")}, current_step=update)
        infos = tr._sample_training_data()
        tr.buffer.prepare_batch_dict()
        stats, _ = tr._train_epochs(lr, clip, beta)
        res = process_episode_info(infos)
        curve.append((res.get("success_percent", float("nan")), res.get("reward_mean", float("nan"))))
        if update % 10 == 9:
            block = np.array(curve[-10:])
            print("updates %3d-%3d  success %.2f  reward %.2f  loss %.4f" % (update - 9, update, np.nanmean(block[:, 0]), np.nanmean(block[:, 1]),
                                                                              float(np.mean(np.array(stats)[:, 2]))), flush=True)
    tr.close(exit_process=False)
    first, last = np.nanmean(np.array(curve[:10])[:, 0]), np.nanmean(np.array(curve[-10:])[:, 0])
    print("RESULT first10 %.3f last10 %.3f" % (first, last))


if __name__ == "__main__":
    main()

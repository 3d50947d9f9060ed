#!/usr/bin/env python
"""Profiling harness for the c3 workload (run under gpurun; numbers printed here are NOT bench values).

  python tools/profile_step.py --kineto gpurun_out/kernels_c3.txt
        one full PPO update under torch.profiler (CUPTI): per-kernel device-time table
  ncu --profile-from-start off ... python tools/profile_step.py --range
        cudaProfilerStart/Stop around `--rollout-steps` rollout steps + `--minibatches` PPO minibatch
        steps, so ncu sees a short, representative slice
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "episodic-transformer-memory-ppo_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def build(workload):
    from bench import load_workload
    from device_feed import SyntheticDeviceFeed
    from trainer import PPOTrainer
    cfg = load_workload(workload)
    env = cfg["environment"]
    os.chdir("/tmp")
    torch.set_num_threads(8)
    tr = PPOTrainer(cfg, run_id="prof", device=torch.device("cuda:0"), workers=[], summary_writer=False)
    tr.device_feed = SyntheticDeviceFeed(cfg["n_workers"], cfg["worker_steps"], tuple(env["obs_shape"]), env["max_episode_steps"],
                                         env.get("min_episode_steps"), seed=7, device="cuda:0")
    return tr, cfg


def update(tr, cfg):
    tr._sample_training_data()
    tr.buffer.prepare_batch_dict()
    tr._train_epochs(cfg["learning_rate_schedule"]["initial"], cfg["clip_range_schedule"]["initial"], cfg["beta_schedule"]["initial"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_minigrid_synthetic")
    ap.add_argument("--kineto", default=None, help="write a per-kernel table of one update to this file")
    ap.add_argument("--range", action="store_true", help="bracket a short slice with cudaProfilerStart/Stop for ncu")
    ap.add_argument("--minibatches", type=int, default=2)
    ap.add_argument("--rollout-steps", type=int, default=2)
    ap.add_argument("--kineto-rollout", default=None, help="per-kernel table of ONE rollout (graphs off) to this file")
    args = ap.parse_args()
    if args.kineto:
        args.kineto = os.path.abspath(args.kineto)
    if args.kineto_rollout:
        args.kineto_rollout = os.path.abspath(args.kineto_rollout)
    tr, cfg = build(args.workload)
    update(tr, cfg)                     # warm-up: allocations, cuDNN heuristics, first-touch
    torch.cuda.synchronize()
    if args.kineto:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            update(tr, cfg)
            torch.cuda.synchronize()
        table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90)
        with open(args.kineto, "w") as f:
            f.write(table)
        print(table[:6000])
    if args.kineto_rollout:
        from torch.profiler import ProfilerActivity, profile
        out = os.path.abspath(args.kineto_rollout) if not os.path.isabs(args.kineto_rollout) else args.kineto_rollout
        tr.use_cuda_graphs = False
        tr._sample_training_data()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr._sample_training_data()
            torch.cuda.synchronize()
        table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=100)
        with open(out, "w") as f:
            f.write(table)
        print(table[:200])
    if args.range:
        tr._sample_training_data()
        tr.buffer.prepare_batch_dict()
        gen = tr.buffer.mini_batch_generator()
        stats = torch.zeros(6, device="cuda:0")
        norms = torch.zeros(tr.model._n_groups + 2, device="cuda:0")
        mbs = [next(gen) for _ in range(args.minibatches + 1)]
        grouping = tr._begin_grouped_attention()
        if grouping is not None:
            tr._group_epoch(mbs, grouping)
        tr._ppo_step(mbs[0], 3e-4, 0.1, 1e-3, stats, norms)         # warm this exact shape
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        tr._rollout_ctx()
        tr._prepare_group(tr._whole)
        with torch.no_grad():
            feed = tr.device_feed
            step = torch.zeros(tr.num_workers, dtype=torch.long, device="cuda:0") + 100
            ep = torch.arange(tr.num_workers, device="cuda:0")
            for t in range(args.rollout_steps):
                tr._device_step(tr._whole, t, (feed.obs(t), step, ep, False))
        for mb in mbs[1:]:
            tr._ppo_step(mb, 3e-4, 0.1, 1e-3, stats, norms)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()

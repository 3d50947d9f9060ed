#!/usr/bin/env python
"""Rollout-only probe of the e2e transport (real env processes) under different grouping / wake-up modes.

    python tools/e2e_probe.py [--workload c3_minigrid_synthetic] [--rollouts 4]

Environment knobs are read by PPOTrainer: TRXL_ROLLOUT_GROUPS, TRXL_SPIN_STEPPING, TRXL_PIN_WORKERS, TRXL_E2E_TRACE."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "episodic-transformer-memory-ppo_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_minigrid_synthetic")
    ap.add_argument("--rollouts", type=int, default=4)
    ap.add_argument("--train", action="store_true", help="also run the optimisation epochs between rollouts")
    ap.add_argument("--profile", action="store_true", help="kineto profile of the last rollout (per-kernel device times)")
    args = ap.parse_args()
    from trainer import PPOTrainer, effective_cpus
    from yaml_parser import YamlParser
    cfg = YamlParser(os.path.join(PKG, "configs", args.workload + ".yaml")).get_config()
    os.chdir(os.environ.get("TMPDIR", "/tmp"))
    tr = PPOTrainer(cfg, run_id="probe", device=torch.device("cuda:0"), summary_writer=False)
    mode = "pipes" if tr._control is None else ("futex" if tr._control.get("futex") is not None else
                                                ("blocking" if tr._control.get("sems") else "spinning"))
    print("[probe] groups=%d mode=%s effective_cpus=%.1f cpu_count=%d" % (len(tr._groups), mode, effective_cpus(), os.cpu_count()), flush=True)
    for r in range(args.rollouts):
        tr.timers = {"rollout": 0.0, "train": 0.0, "env": 0.0}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.profile and r == args.rollouts - 1:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                tr._sample_training_data()
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70), flush=True)
        else:
            tr._sample_training_data()
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tr.buffer.prepare_batch_dict()
        if args.train:
            tr._train_epochs(3e-4, 0.1, 1e-3)
            torch.cuda.synchronize()
        print("[probe] rollout %d: %.1f ms (%.3f ms/step), env-phase share %.1f ms, train %.1f ms"
              % (r, dt * 1e3, dt * 1e3 / cfg["worker_steps"], tr.timers["env"] * 1e3, tr.timers["train"] * 1e3), flush=True)
    tr.close(exit_process=False)


if __name__ == "__main__":
    main()

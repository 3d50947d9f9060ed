#!/usr/bin/env python
"""Per-phase latency of the fused rollout forward (cluster 0's SM clock at every phase boundary).

    TRXL_RF_TRACE=0 python tools/rf_trace.py [--workload c3_minigrid_synthetic] [--steps 3]

Runs one rollout to fill the episode tables, then `--steps` eager rollout steps with the trace switched on; the library
prints the phase table to stderr (csrc/rollout_fused.cu).  Debugging aid: the numbers are per-phase clocks of ONE cluster.
"""
import argparse
import os
import sys

os.environ.setdefault("TRXL_RF_TRACE", "0")
os.environ.setdefault("TRXL_CONV_TRACE", "0")           # must exist before the first launch for the library to arm the trace
os.environ["TRXL_NO_GRAPHS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402
from profile_step import build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3_minigrid_synthetic")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    tr, cfg = build(args.workload)
    tr._sample_training_data()                        # fills the episode tables; trace off
    torch.cuda.synchronize()
    traced = set(range(300, 300 + args.steps))
    inner = tr._step_via_graph

    def step(mode, grp, t, src):
        os.environ["TRXL_RF_TRACE"] = "1" if t in traced else "0"
        os.environ["TRXL_CONV_TRACE"] = "1" if t in traced else "0"
        return inner(mode, grp, t, src)

    tr._step_via_graph = step
    tr._sample_training_data()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()

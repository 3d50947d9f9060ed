#!/usr/bin/env python
"""Per-parameter gradient error of one full-size c3 optimiser step: CUDA vs the fp32 oracle vs the float64 oracle."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "episodic-transformer-memory-ppo_b200")):
    sys.path.insert(0, p)
os.chdir(os.environ.get("TMPDIR", "/tmp"))
from test_gpu_timed_path import c3_minibatch_errors  # noqa: E402
for safe in (True, False):
    print("=== ReLU-safe samples" if safe else "=== first shuffled minibatch as is", flush=True)
    (stats, s32, s64), fwd, _ = c3_minibatch_errors(verbose=True, safe=safe)
    print("forward:", fwd)
    print("stats cuda", stats, "\nstats o32 ", s32, "\nstats o64 ", s64, flush=True)

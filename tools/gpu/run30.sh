#!/bin/bash
# per-phase latency of the fused rollout kernel
cd /root/repo
python tools/rf_trace.py --steps 2 > gpurun_out/rf_trace.txt 2>&1
tail -60 gpurun_out/rf_trace.txt

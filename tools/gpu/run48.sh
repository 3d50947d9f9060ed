#!/bin/bash
# e2e rollout vs number of worker groups, final build
cd /root/repo
for g in 6 8 11 16; do
  TRXL_ROLLOUT_GROUPS=$g python tools/e2e_probe.py --rollouts 5 > gpurun_out/r48_groups$g.txt 2>&1; echo "groups $g:"; tail -2 gpurun_out/r48_groups$g.txt
done

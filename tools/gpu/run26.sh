cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/e2e_probe.py --rollouts 6 2>&1 | tail -4
TRXL_E2E_TRACE=1 timeout 300 python tools/e2e_probe.py --rollouts 4 2>&1 | tail -2

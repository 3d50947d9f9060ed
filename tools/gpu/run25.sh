cd $GRAFT_REPO_ROOT
nproc
TRXL_E2E_TRACE=1 timeout 300 python tools/e2e_probe.py --rollouts 5 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu/run23.sh

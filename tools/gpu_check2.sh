cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_full_horizon_parity.py -q -s -k "every_solver_block or wave" 2>&1 | grep -v "^$" | tail -14

cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "debris or retyp or drop_in" 2>&1 | grep -A14 "Mismatch\|passed\|failed" | head -30
MPCB_LIB=/root/repo/mpc_arpo_project_b200/lib/libmpcb_gp.so python tools/debris_bench.py 2>&1 | grep -v "^$" | awk 'NR%3!=1' | tail -8

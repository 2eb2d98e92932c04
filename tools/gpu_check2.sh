cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "debris or retyp or drop_in" 2>&1 | tail -8
python tools/debris_bench.py 2>&1 | tail -4
MPCB_LIB=/root/repo/mpc_arpo_project_b200/lib/libmpcb_gp.so python tools/debris_bench.py 2>&1 | grep -v "^$" | tail -12

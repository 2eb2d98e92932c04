cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_full_horizon_parity.py -q -x -k "full_size or reproducible or retyp or wave or every_solver_block or pipeline" 2>&1 | tail -4
for se in 1 4 8; do echo -n "sync_every=$se: "; MPCB_SYNC_EVERY=$se BENCH_VERBOSE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --parity-lanes 0 2>&1 | grep "\[bench\] step" | head -8 | awk '{s+=$6; printf "%s ", $6} END {print " avg", s/NR}'; done
for se in 1 4; do echo -n "64k sync_every=$se: "; MPCB_SYNC_EVERY=$se BENCH_VERBOSE=1 timeout 900 python bench.py --workload config2 --lanes 65536 --steps 3 --warmup 2 --no-cpu-baseline --parity-lanes 0 2>&1 | grep "\[bench\] step" | head -3 | awk '{printf "%s ", $6}'; echo; done

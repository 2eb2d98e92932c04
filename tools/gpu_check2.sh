cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_full_horizon_parity.py -q -x -k "full_size or reproducible or retyp or wave" 2>&1 | tail -4
for sl in 0 1000 2000 4000; do echo -n "slice=$sl: "; MPCB_SLICE_ITERS=$sl BENCH_VERBOSE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --parity-lanes 0 2>&1 | grep "\[bench\] step" | head -8 | awk '{s+=$6; printf "%s ", $6} END {print " avg", s/NR}'; done

cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_kf_variant.py -q -x -k "continuous or contC or every_solver_block or reproducible" 2>&1 | tail -3
BENCH_VERBOSE=1 timeout 900 python bench.py --workload config3 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "\[bench\] step" | head -3 | awk '{printf "%s ", $6}'; echo
BENCH_VERBOSE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --parity-lanes 0 2>&1 | grep "\[bench\] step" | head -6 | awk '{printf "%s ", $6}'; echo
BENCH_VERBOSE=1 timeout 900 python bench.py --workload config2 --lanes 65536 --steps 3 --warmup 3 --no-cpu-baseline --parity-lanes 0 2>&1 | grep "\[bench\] step" | head -3 | awk '{printf "%s ", $6}'; echo

cd $GRAFT_REPO_ROOT
for rb in 2500 2960 3300 3600; do echo -n "rb=$rb: "; MPCB_RESUME_BELOW=$rb BENCH_VERBOSE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --parity-lanes 0 2>&1 | grep "\[bench\] step" | head -6 | awk '{s+=$6; printf "%s ", $6} END {print " avg", s/NR}'; done

# usage: ab.sh libA libB ... ; prints per-step device-leg ms for each library, alternating twice
for i in 1 2; do for lib in "$@"; do echo -n "$lib: "; BENCH_VERBOSE=1 MPCB_LIB=/root/repo/mpc_arpo_project_b200/lib/$lib timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep "\[bench\]" | head -4 | awk '{printf "%s ", $3}'; echo; done; done

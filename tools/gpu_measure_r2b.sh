cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -k "retyp or reproducible or full_size or every_solver_block or monte_carlo" 2>&1 | tail -5
b() { name=$1; shift; BENCH_VERBOSE=1 timeout 1200 python bench.py "$@" > $O/r2_bench_$name.json 2> $O/r2_bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/r2_bench_$name.json'));print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config'].get('flip_lanes'), d.get('parity'))" 2>&1 | tail -1)"; }
b config2_64k --workload config2 --lanes 65536 --steps 3 --warmup 3 --no-cpu-baseline
b config2_quiet_64k --workload config2_quiet --lanes 65536 --steps 3 --warmup 3 --no-cpu-baseline
b config2_quiet --workload config2_quiet --steps 5 --warmup 3 --no-cpu-baseline
b config4 --workload config4 --steps 3 --warmup 3

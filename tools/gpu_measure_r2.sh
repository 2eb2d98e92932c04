# Round-2 measurement set: GPU tests, bench lines (>= 3 warm-up steps) for BASELINE configs 1-5 + large-batch lines.
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rA -s > $O/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -n "engine tables\|control (\|FAILED\|passed\|failed" $O/r2c_pytest_gpu.log | tail -20
b() { name=$1; shift; BENCH_VERBOSE=1 timeout 1200 python bench.py "$@" > $O/r2_bench_$name.json 2> $O/r2_bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/r2_bench_$name.json'));print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; }
b config2
b config3 --workload config3 --steps 3 --warmup 3
b config5 --workload config5 --steps 2 --warmup 3
b config1 --workload config1 --steps 2 --warmup 3
b reference --impl reference --steps 3 --warmup 1

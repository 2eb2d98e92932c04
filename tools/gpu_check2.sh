cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "debris or drop_in or reproducible" 2>&1 | tail -3
BENCH_VERBOSE=1 timeout 1500 python bench.py --workload config1 --steps 2 --warmup 3 > gpurun_out/r2_bench_config1.json 2> gpurun_out/r2_bench_config1.err; tail -3 gpurun_out/r2_bench_config1.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench_config1.json'));print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'], d['cpu_baseline']['value'])"

cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_full_horizon_parity.py tests/test_kf_variant.py -q -x -k "continuous or plant or contC" 2>&1 | tail -4
BENCH_VERBOSE=1 timeout 900 python bench.py --workload config3 --steps 3 --warmup 3 > gpurun_out/r2_bench_config3.json 2> gpurun_out/r2_bench_config3.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench_config3.json'));print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])"

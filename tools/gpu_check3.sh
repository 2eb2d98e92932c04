cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "debris or drop_in or reproducible" 2>&1 | tail -3
python tools/debris_bench.py 2>&1 | grep -v "^$" | tail -3
BENCH_VERBOSE=1 timeout 1500 python bench.py --workload config1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/exp_c1.json 2> gpurun_out/exp_c1.err; python -c "import json;d=json.load(open('gpurun_out/exp_c1.json'));print(round(d['value']), d['ms_per_step'])"

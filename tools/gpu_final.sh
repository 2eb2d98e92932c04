# Final check of the round: smoke, the whole GPU suite, the default bench line and the reference arm, as the driver runs them.
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2700 python -m pytest tests -m gpu -q -x > $O/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_final_pytest_gpu.log
timeout 900 python bench.py > $O/r2_bench_config2.json 2> $O/r2_bench_config2.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('$O/r2_bench_config2.json'));print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline']['value'], d['parity']['exact_frac'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/r2_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_launches_config2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-lanes 0 > $O/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"

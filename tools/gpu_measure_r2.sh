# Round-2 measurement set: bench lines (>= 3 warm-up steps) for BASELINE configs 2-5 + large-batch lines, launch list, ncu captures.
cd $GRAFT_REPO_ROOT
O=gpurun_out
b() { name=$1; shift; timeout 900 python bench.py "$@" > $O/r2_bench_$name.json 2> $O/r2_bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/r2_bench_$name.json'));print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))" 2>&1 | tail -1)"; }
b config2 
b config2_quiet --workload config2_quiet --steps 5 --warmup 3 --no-cpu-baseline
b config2_64k --workload config2 --lanes 65536 --steps 3 --warmup 3 --no-cpu-baseline
b config2_quiet_64k --workload config2_quiet --lanes 65536 --steps 3 --warmup 3 --no-cpu-baseline
b config3 --workload config3 --steps 3 --warmup 3
b config4 --workload config4 --steps 3 --warmup 3
b config5 --workload config5 --steps 2 --warmup 3
b reference --impl reference --steps 3 --warmup 1
# launch list of the default command (short run) and full captures of the two solver kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_config2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-lanes 0 > $O/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:team_kernel -s 1 -c 1 -o $O/r2_prof_team_config2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-lanes 0 > $O/r2_ncu_team.log 2>&1; echo "ncu team rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_config2_64k.csv python bench.py --workload config2 --lanes 65536 --steps 1 --warmup 1 --no-cpu-baseline --parity-lanes 0 > $O/r2_ncu_launches64k.log 2>&1; echo "launch list 64k rc=$?"

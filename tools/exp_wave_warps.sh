# A/B of wave-kernel CTA sizes (8 vs 4 warps) at large and BASELINE batch sizes; prints ms per step (device leg)
cd $GRAFT_REPO_ROOT
L=/root/repo/mpc_arpo_project_b200/lib
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" BENCH_VERBOSE=1 timeout 400 python bench.py --no-cpu-baseline --parity-lanes 0 $BARGS > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  echo "$name: $(grep '\[bench\] step' gpurun_out/exp_$name.err | awk '{printf "%s ", $6}')  value=$(python -c "import json;print(round(json.load(open('gpurun_out/exp_$name.json'))['value']))")"
}
BARGS="--workload config2 --lanes 65536 --steps 2 --warmup 1"
run w8_64k MPCB_LIB=$L/libmpcb.so
run w4_64k MPCB_LIB=$L/libmpcb_w4.so
BARGS="--workload config2 --steps 4 --warmup 3"
run team_4k MPCB_LIB=$L/libmpcb.so
run w8_4k MPCB_LIB=$L/libmpcb.so MPCB_WAVE_MIN_LANES=4096
run w4_4k MPCB_LIB=$L/libmpcb_w4.so MPCB_WAVE_MIN_LANES=4096
run w4_4k_rb600 MPCB_LIB=$L/libmpcb_w4.so MPCB_WAVE_MIN_LANES=4096 MPCB_RESUME_BELOW=600
run w4_4k_rb1500 MPCB_LIB=$L/libmpcb_w4.so MPCB_WAVE_MIN_LANES=4096 MPCB_RESUME_BELOW=1500
BARGS="--workload config2 --lanes 16384 --steps 3 --warmup 2"
run team_16k MPCB_LIB=$L/libmpcb.so MPCB_SOLVER=team
run w8_16k MPCB_LIB=$L/libmpcb.so
run w4_16k MPCB_LIB=$L/libmpcb_w4.so

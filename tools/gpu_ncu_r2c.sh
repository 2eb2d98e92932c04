# DRAM traffic of the two solver kernels of the default command (BASELINE config 2, 4096 lanes): one mid-run wave round and the hand-over launch
cd $GRAFT_REPO_ROOT
O=gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --parity-lanes 0"
ncu --set full --clock-control none --import-source on -k regex:admm_wave -s 250 -c 1 -o $O/r2_prof_wave4_config2 -f python bench.py $ARGS > $O/r2_ncu_wave4.log 2>&1; echo "ncu wave4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:team_kernel -s 1 -c 1 -o $O/r2_prof_team_resume_config2 -f python bench.py $ARGS > $O/r2_ncu_team_resume.log 2>&1; echo "ncu team resume rc=$?"

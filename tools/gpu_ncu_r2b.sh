# list mode of the continuous simulator (config 3 at 8192 lanes): launch list + full captures of one team_kernel (list mode) and one post_kernel launch
cd $GRAFT_REPO_ROOT
O=gpurun_out
ARGS="--workload config3 --lanes 8192 --steps 1 --warmup 1 --no-cpu-baseline --parity-lanes 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r2_launches_config3_8k.csv python bench.py $ARGS > $O/r2_ncu_launches_c3.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:team_kernel -s 40 -c 1 -o $O/r2_prof_team_listmode_config3 -f python bench.py $ARGS > $O/r2_ncu_team_list.log 2>&1; echo "ncu team list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 40 -c 1 -o $O/r2_prof_post_config3 -f python bench.py $ARGS > $O/r2_ncu_post.log 2>&1; echo "ncu post rc=$?"

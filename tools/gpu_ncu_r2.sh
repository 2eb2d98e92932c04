cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/ncu_debris.py > $O/ncu_debris_plain.log 2>&1; echo "plain rc=$?"; tail -1 $O/ncu_debris_plain.log
ncu --set full --clock-control none --import-source on -k regex:generic_lane -s 1 -c 1 -o $O/r2_prof_generic_config1 -f python tools/ncu_debris.py > $O/r2_ncu_generic.log 2>&1; echo "ncu generic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:admm_wave -s 60 -c 1 -o $O/r2_prof_wave_config2_64k -f python bench.py --workload config2 --lanes 65536 --steps 1 --warmup 1 --no-cpu-baseline --parity-lanes 0 > $O/r2_ncu_wave.log 2>&1; echo "ncu wave rc=$?"
ls -la $O/*.ncu-rep | tail -3

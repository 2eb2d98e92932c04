set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rA -s > gpurun_out/r2b_pytest_gpu_all.log 2>&1; echo "pytest rc=$?"
grep -n "engine tables\|control (\|PASSED\|FAILED\|passed\|failed" gpurun_out/r2b_pytest_gpu_all.log | tail -80

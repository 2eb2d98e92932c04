cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_2gpu.out 2> gpurun_out/r2_2gpu.err; echo "rc=$?"
wc -c gpurun_out/r2_2gpu.out gpurun_out/r2_2gpu.err
tail -20 gpurun_out/r2_2gpu.err
head -c 600 gpurun_out/r2_2gpu.out

# weak-scaling lines as the driver launches them (torchrun, one rank per GPU); usage: gpurun --gpus N -- bash tools/gpu_scale.sh "N [N2 ...]"
cd $GRAFT_REPO_ROOT
O=gpurun_out
for n in ${1:-8 4}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > $O/r2_scale_n$n.json 2> $O/r2_scale_n$n.err; echo "n=$n rc=$?"
  python -c "import json;l=[x for x in open('$O/r2_scale_n$n.json') if x.startswith('{')];d=json.loads(l[0]);print(len(l), d['n_gpus'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'])"
done

# usage: gpu_tp_bench.sh <N> <workload> ; N ranks form one tensor-parallel group (N = 1: single-GPU persistent engine)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1; WL=$2
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/tp_${WL}_n1.json 2> gpurun_out/tp_${WL}_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --tp --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/tp_${WL}_n$N.json 2> gpurun_out/tp_${WL}_n$N.err
fi
tail -c 600 gpurun_out/tp_${WL}_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/tp_${WL}_n$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','scaling','ms_per_step')}, d['config']['parallelism'], d['whole_step']['us_per_token'], d['whole_step']['GBps'], d['tokens_tail'])
PY

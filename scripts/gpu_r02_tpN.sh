# usage: gpu_r02_tpN.sh <N>: the driver's scaling command at N GPUs (70B INT4, TP strong scaling with in-bench parity)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_tp${N}_bench.json 2> gpurun_out/r02_tp${N}_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_tp${N}_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','scaling','ms_per_step')}, d['roofline']['frac'], d['tp_tokens_equal'], d['tp_logits_rel_err'], d['strong_scaling'])
PY
tail -4 gpurun_out/r02_tp${N}_bench.err

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for N in 8 4; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_ll_tp${N}_bench.json 2> gpurun_out/r02_ll_tp${N}_bench.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_ll_tp${N}_bench.json') if l.startswith('{')][-1])
print($N, round(d['value'],1), d['roofline']['frac'], d['tp_tokens_equal'], d['tp_logits_rel_err'], d['strong_scaling']['speedup'])
PY
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline --workload llama70b-int4-ctx4096-decode64 ) > gpurun_out/r02_ll_cfg5_tp8_bench.json 2> gpurun_out/r02_ll_cfg5_tp8_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_ll_cfg5_tp8_bench.json') if l.startswith('{')][-1])
print('cfg5', round(d['value'],1), d['roofline']['frac'], d['tp_tokens_equal'], d['strong_scaling'])
PY

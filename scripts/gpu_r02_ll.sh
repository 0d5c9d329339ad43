cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1
timeout 1500 python -m pytest tests/test_gpu_tp.py -m gpu -q -x 2>&1 | tail -6
for mode in ll p2p barrier; do
  TURBOINFER_B200_TP_REDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$mode', round(d['value'],1), d['tp_tokens_equal'], d['tp_logits_rel_err'], d['strong_scaling']['speedup'])"
done

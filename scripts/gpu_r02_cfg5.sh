# BASELINE.json configs[4] as written on 8 GPUs: Llama-2-70B shape, INT4, TP 8, 4k-context paged KV cache, batch 1 (bench.py) --
# and the default TP workload (decode-64, short context) with the final build, both with the in-bench parity check.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=8
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --workload llama70b-int4-ctx4096-decode64 ) > gpurun_out/r02_cfg5_tp8_bench.json 2> gpurun_out/r02_cfg5_tp8_bench.err
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_final_tp8_bench.json 2> gpurun_out/r02_final_tp8_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_cfg5_tp8_bench.json','gpurun_out/r02_final_tp8_bench.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','n_gpus','scaling','ms_per_step')}, d['roofline']['frac'], d.get('tp_tokens_equal'), d.get('tp_logits_rel_err'), d.get('strong_scaling'), d.get('prefill'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -5 gpurun_out/r02_cfg5_tp8_bench.err; tail -4 gpurun_out/r02_final_tp8_bench.err

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_fullwidth.py -m gpu -q 2>&1 | tail -8
run() { python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --workload llama7b-int4-prefill2048-decode256 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), d.get('prefill_ms'), d['tokens_tail'], d.get('e2e'))"; }
run tc
TURBOINFER_B200_PREFILL_ATTN=fp32 run fp32

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py tests/test_gpu_batch.py tests/test_gpu_prefill.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python scripts/timeline.py llama7b 3 260 2>&1 | tee gpurun_out/r02g_timeline.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02g_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['tokens_tail'])
PY
tail -3 gpurun_out/r02g_bench.err
for w in tinyllama-int4-decode512 llama7b-int8-decode256; do timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['config']['workload'], round(d['value'],1), round(d['roofline']['frac'],3), d['tokens_tail'])"; done

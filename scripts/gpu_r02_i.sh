cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batch.py tests/test_gpu_prefill.py tests/test_gpu_fullwidth.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python scripts/timeline.py llama7b 3 260 2>&1 | tee gpurun_out/r02i_timeline.txt
timeout 200 python scripts/timeline.py llama7b 3 16 2>&1 | grep attn
timeout 200 python scripts/timeline.py llama7b 3 1000 2>&1 | grep attn
timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['tokens_tail'])
PY
tail -3 gpurun_out/r02i_bench.err

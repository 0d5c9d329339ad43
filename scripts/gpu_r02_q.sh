cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_fullwidth.py -m gpu -q 2>&1 | tail -8
run() { python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --workload llama7b-int4-prefill2048-decode256 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), d.get('prefill'), d['tokens_tail'])"; }
run h3
TURBOINFER_B200_PREFILL_ATTN=tf32 run tf32
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'causal' -c 10 --csv --log-file gpurun_out/r02q_attn.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline --workload llama7b-int4-prefill2048-decode256 > /dev/null 2>&1
grep -c causal gpurun_out/r02q_attn.csv; tail -3 gpurun_out/r02q_attn.csv | cut -c1-400
python bench.py --steps 2 --warmup 2 --no-extras --no-cpu-baseline --workload tinyllama-int4-decode512 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('tiny', round(d['value'],1), d['tokens_tail'])"

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline $2 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), d['roofline']['frac'], d['tokens_tail'])"; }
for c in 48 64 80 96; do TURBOINFER_B200_ATTN_CHUNK=$c run chunk$c; done
run tiny_base "--workload tinyllama-int4-decode512"
for c in 64 80 128; do TURBOINFER_B200_ATTN_CHUNK=$c run tiny_chunk$c "--workload tinyllama-int4-decode512"; done
run pf_base "--workload llama7b-int4-prefill2048-decode256"
for c in 80 160; do TURBOINFER_B200_ATTN_CHUNK=$c run pf_chunk$c "--workload llama7b-int4-prefill2048-decode256"; done

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), d['roofline']['frac'], d['tokens_tail'])"; }
run base
for c in 80 128 160 200; do TURBOINFER_B200_ATTN_CHUNK=$c run chunk$c; done
bb() { python scripts/bench_batch.py --shape llama7b --qtype int4 --batch 32 --new 128 --reps 2 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), {k:v for k,v in d.items() if 'ms' in k})"; }
bb batch_base
for k in 4 8 12 16; do TURBOINFER_B200_SPLIT_KSTEPS=$k bb split$k; done

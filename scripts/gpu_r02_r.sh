cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_batch.py tests/test_gpu_beam.py tests/test_gpu_fullwidth.py tests/test_gpu_decode.py tests/test_gpu_sampling.py tests/test_host_cpp.py -m gpu -q 2>&1 | tail -8
bb() { python scripts/bench_batch.py --shape llama7b --qtype int4 --batch $2 --new 256 --reps 2 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), {k:v for k,v in d.items() if 'ms' in k})"; }
bb pdl32 32
TURBOINFER_B200_BATCH_PDL=0 bb nopdl32 32
bb pdl8 8
TURBOINFER_B200_BATCH_PDL=0 bb nopdl8 8

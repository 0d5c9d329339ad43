cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_gemm.py tests/test_gpu_fullwidth.py tests/test_gpu_beam.py -m gpu -q 2>&1 | tail -4
bb() { python scripts/bench_batch.py --shape llama7b --qtype int4 --batch $2 --new 256 --reps 2 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), {k:v for k,v in d.items() if 'ms' in k})"; }
bb tiles2_b32 32
TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_T1.so bb tiles1_b32 32
TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_T3.so bb tiles3_b32 32
bb tiles2_b8 8
TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_T1.so bb tiles1_b8 8

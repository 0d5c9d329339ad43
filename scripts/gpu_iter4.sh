cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for fl in 0 1 2 3; do
for w in llama7b-int4-decode256 tinyllama-int4-decode512; do
TURBOINFER_B200_DBG_FLAGS=$fl timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('flags $fl', '$w', d['value'], d['e2e']['value'], d['whole_step']['us_per_token'], d['whole_step']['frac_of_measured_peak'], d['tokens_tail'])"; tail -3 gpurun_out/bench_$w.err
done; done
timeout 250 python scripts/timeline.py llama7b 2 16 2>&1 | tail -16

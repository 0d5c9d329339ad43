cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for w in tinyllama-int4-decode512 llama7b-int4-decode256; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; python -c "
import json; d=json.load(open('gpurun_out/q_$w.json')); print('$w', round(d['value'],1), round(d['whole_step']['us_per_token'],1), round(d['whole_step']['frac_of_measured_peak'],3), d['tokens_tail'])"; tail -3 gpurun_out/q_$w.err
done

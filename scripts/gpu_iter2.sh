cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 ./scripts/microbench3 2>&1 | head -16
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_gpu_ops.py tests/test_gpu_gemm.py -m gpu -q -x --timeout 600 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_prefill.py -m gpu -q -x --timeout 600 2>&1 | tail -8
timeout 250 python scripts/timeline.py tinyllama 2 16 2>&1 | tail -14
timeout 250 python scripts/timeline.py llama7b 2 16 2>&1 | tail -14
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tinyllama.json 2> gpurun_out/bench_tinyllama.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tinyllama.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['whole_step']['us_per_token'], d['whole_step']['frac_of_measured_peak'], d['roofline']['frac'], d['tokens_tail'])"; tail -5 gpurun_out/bench_tinyllama.err
timeout 600 python bench.py --workload llama7b-int4-decode256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_7b_int4.json 2> gpurun_out/bench_7b_int4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_7b_int4.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['whole_step']['us_per_token'], d['whole_step']['frac_of_measured_peak'], d['roofline']['frac'], d['tokens_tail'])"; tail -5 gpurun_out/bench_7b_int4.err

cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5
timeout 250 python scripts/timeline.py llama7b 2 16 2>&1 | tail -15
timeout 600 python bench.py --workload llama7b-int4-decode256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_7b_int4.json 2> gpurun_out/bench_7b_int4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_7b_int4.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['whole_step']['us_per_token'], d['roofline'], d['tokens_tail'])"; tail -5 gpurun_out/bench_7b_int4.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tinyllama.json 2> gpurun_out/bench_tinyllama.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tinyllama.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['whole_step']['us_per_token'], d['roofline']['frac'], d['tokens_tail'])"; tail -5 gpurun_out/bench_tinyllama.err

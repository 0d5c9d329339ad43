cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --workload llama7b-int4-prefill2048-decode256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_7b_prefill.json 2> gpurun_out/bench_7b_prefill.err; tail -3 gpurun_out/bench_7b_prefill.err; python -c "
import json; d=json.load(open('gpurun_out/bench_7b_prefill.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['prefill'], d['whole_step']['us_per_token'], d['roofline']['frac'], d['tokens_tail'])"

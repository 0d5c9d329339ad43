cd $GRAFT_REPO_ROOT
python -c "
from turboinfer_b200 import build as b
print(b.build_host_test('tests/cpp/test_host_api'))"
./tests/cpp/test_host_api gpu 2>&1 | tail -20
timeout 600 python bench.py --workload llama7b-int4-decode256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_7b_int4.json 2> gpurun_out/bench_7b_int4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_7b_int4.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['whole_step']['us_per_token'], d['roofline']['frac'], d['tokens_tail'])"; tail -5 gpurun_out/bench_7b_int4.err

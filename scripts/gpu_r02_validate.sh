# What the driver runs at round end, in one call: the GPU test suite, smoke(), both bench arms at N = 1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_validate_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/r02_validate_bench_ref.json 2> gpurun_out/r02_validate_bench_ref.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/r02_validate_bench.json 2> gpurun_out/r02_validate_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_validate_bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','steps','warmup','ms_per_step','e2e','gpu_launches','clocks')}, d['roofline']['frac'], d['roofline']['traffic'])
print({k:(v.get('value'), v.get('prefill_ms')) for k,v in d.get('extras',{}).items()})
r=json.loads([l for l in open('gpurun_out/r02_validate_bench_ref.json') if l.startswith('{')][-1])
print({k:r.get(k) for k in ('impl','value','ms_per_step','steps')})
PY
tail -4 gpurun_out/r02_validate_bench.err

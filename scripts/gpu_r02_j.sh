cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02j_pytest.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02j_bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline'], d.get('per_kernel'), d.get('extras'))
PY
tail -3 gpurun_out/r02j_bench.err
TURBOINFER_B200_PDL=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02j_bench_nopdl.json 2>/dev/null; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02j_bench_nopdl.json') if l.startswith('{')][-1])
print('nopdl', {k:d.get(k) for k in ('value','ms_per_step')}, d.get('per_kernel'))
PY

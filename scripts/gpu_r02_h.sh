cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02h_plain.json 2> gpurun_out/r02h_plain.err && python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h_plain.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['tokens_tail'])
PY
$CMD > gpurun_out/r02h_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mega_decode_kernel -s 7 -c 1 -o gpurun_out/r02h_prof_mega $CMD > gpurun_out/r02h_ncu.log 2>&1
tail -3 gpurun_out/r02h_ncu.log | cut -c1-300
ls -la gpurun_out/r02h*

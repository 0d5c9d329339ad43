cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --workload llama7b-int4-prefill2048-decode256 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value'],1), d.get('prefill'))"; }
run tc
TURBOINFER_B200_PREFILL_ATTN=fp32 run fp32
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'causal|gemm_i8_tc_kernel|rmsnorm_digits_kernel|rope_kv_kernel' -c 60 --csv --log-file gpurun_out/r02p_prefill_launches.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline --workload llama7b-int4-prefill2048-decode256 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02p_prefill_launches.csv')))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
            agg[d['Kernel Name'][:50]][0]+=1; agg[d['Kernel Name'][:50]][1]+=v
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v[1]/v[0]:9.1f} us avg x{v[0]:3d}  {k}")
PY

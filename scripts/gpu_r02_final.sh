# Final measurement pass of round 2 (one GPU): bench both arms, ncu launch list, ncu --set full of the decode launch, the tcgen05
# GEMM and the tensor-core prefill attention, and the per-phase timeline.  Each ncu pass runs only after the same command exited 0.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err
CMD="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_final_plain.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mega_decode_kernel -s 7 -c 1 -o gpurun_out/r02_prof_mega -f $CMD > gpurun_out/r02_ncu_mega.log 2>&1
python scripts/gemm_one.py > gpurun_out/r02_gemm_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_tc_kernel -s 2 -c 1 -o gpurun_out/r02_prof_gemm -f python scripts/gemm_one.py > gpurun_out/r02_ncu_gemm.log 2>&1
PF="python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline --workload llama7b-int4-prefill2048-decode256"
$PF > gpurun_out/r02_final_prefill.json 2>/dev/null && \
ncu --set full --clock-control none --import-source on -k regex:causal_attention_tc -s 3 -c 1 -o gpurun_out/r02_prof_attn_tc -f $PF > gpurun_out/r02_ncu_attn.log 2>&1
python scripts/timeline.py llama7b 3 260 > gpurun_out/r02_final_timeline_t260.txt 2>&1
python scripts/timeline.py llama7b 3 16 > gpurun_out/r02_final_timeline_t16.txt 2>&1
python scripts/timeline.py tinyllama 3 16 > gpurun_out/r02_final_timeline_tinyllama.txt 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_final_bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}, d['roofline']['frac'])
r=json.loads([l for l in open('gpurun_out/r02_final_bench_ref.json') if l.startswith('{')][-1])
print({k:r.get(k) for k in ('value','ms_per_step','cpu_baseline')})
PY
ls -la gpurun_out/r02_prof_* gpurun_out/r02_launches.csv

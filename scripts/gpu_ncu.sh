# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list of the hot-path kernels, then one full
# capture of the decode launch of the persistent kernel (launches alternate prefill / decode: index 7 is a decode)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mega_decode|gemv_kernel|attn_" -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mega_decode_kernel -s 7 -c 1 -o gpurun_out/prof_mega $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
timeout 250 python scripts/timeline.py tinyllama 2 16 > gpurun_out/timeline_tinyllama.txt 2>&1
timeout 250 python scripts/timeline.py llama7b 2 16 > gpurun_out/timeline_llama7b.txt 2>&1
timeout 250 python scripts/timeline.py llama7b 2 512 > gpurun_out/timeline_llama7b_t512.txt 2>&1
ls -la gpurun_out

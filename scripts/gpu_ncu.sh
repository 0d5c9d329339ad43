# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list, then one full capture of the GEMV
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 400 -c 5 -o gpurun_out/prof_gemv $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

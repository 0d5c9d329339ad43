set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/gemm_one.py > gpurun_out/gemm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_tc_kernel -s 2 -c 1 -o gpurun_out/prof_gemm python scripts/gemm_one.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-200

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_beam.py tests/test_host_cpp.py -m gpu -q 2>&1 | tail -6
python scripts/bench_sampled.py llama7b 256 2>/dev/null | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sample_kernel|beam_expand" -s 30 -c 6 --csv python scripts/bench_sampled.py llama7b 64 2>/dev/null | grep -E "sample_kernel|beam_expand" | awk -F"\",\"" "{print \$5, \$NF}" | cut -c1-30,60-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sample_kernel|beam_expand" -s 200 -c 6 --csv python scripts/bench_sampled.py llama7b 64 2>/dev/null | grep -E "sample_kernel|beam_expand" | awk -F"\",\"" "{print \$5, \$NF}" | cut -c1-30,60-200

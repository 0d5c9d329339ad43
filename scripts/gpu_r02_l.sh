cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_beam.py tests/test_host_cpp.py tests/test_gpu_sampling.py tests/test_gpu_batch.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02l_pytest.txt

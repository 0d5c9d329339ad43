cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_host_cpp.py -m gpu -q -x --timeout 300 2>&1 | tail -15

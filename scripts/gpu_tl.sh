cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_prefill.py -m gpu -q -x --timeout 600 2>&1 | tail -15

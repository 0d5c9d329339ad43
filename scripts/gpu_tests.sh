set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import turboinfer_b200 as t; t.init(0); print(t.device_info())"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40

cd $GRAFT_REPO_ROOT
timeout 250 python scripts/timeline.py tinyllama 2 16 2>&1 | tail -16
timeout 250 python scripts/timeline.py llama7b 2 16 2>&1 | tail -16

cd $GRAFT_REPO_ROOT
TURBOINFER_B200_DBG_NOMATH=2 timeout 250 python scripts/timeline.py llama7b 2 16 2>&1 | tail -15

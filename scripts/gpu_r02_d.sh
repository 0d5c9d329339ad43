cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for f in 0 8 16 24; do
  echo "=== DBG_FLAGS=$f"
  TURBOINFER_B200_DBG_FLAGS=$f timeout 200 python scripts/timeline.py llama7b 2 260 2>&1 | sed -n 7,13p
done | tee gpurun_out/r02d_loop_experiments.txt

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 scripts/timeline_tp.py llama70b 3 16 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tee gpurun_out/r02_timeline_tp${N}_llama70b.txt | tail -40

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x 2>&1 | tail -15
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_tp2_bench.json 2> gpurun_out/r02_tp2_bench.err
tail -c 2500 gpurun_out/r02_tp2_bench.json; tail -8 gpurun_out/r02_tp2_bench.err

# BASELINE.json configs[4], the batched half: Llama-2-70B shape, INT4, TP 8, 4k-context paged KV cache (fp32), batch 16
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 scripts/bench_batch.py --shape llama70b --qtype int4 --batch 16 --prompt 4032 --new 64 --reps 1 --no-warmup --no-single --tp ) > gpurun_out/r02_cfg5_batch16_tp8.json 2> gpurun_out/r02_cfg5_batch16_tp8.err
grep '^{' gpurun_out/r02_cfg5_batch16_tp8.json | tail -1 | cut -c1-1500
tail -6 gpurun_out/r02_cfg5_batch16_tp8.err
nvidia-smi --query-gpu=memory.used --format=csv | head -3

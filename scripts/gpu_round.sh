# one GPU call: tests, smoke, bench, then ncu launch list + one full capture of the dominant kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tinyllama.json 2> gpurun_out/bench_tinyllama.err; tail -c 3000 gpurun_out/bench_tinyllama.json; tail -5 gpurun_out/bench_tinyllama.err
python bench.py --workload llama7b-int4-decode256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_7b_int4.json 2> gpurun_out/bench_7b_int4.err; tail -c 3000 gpurun_out/bench_7b_int4.json; tail -5 gpurun_out/bench_7b_int4.err
python bench.py --workload llama7b-int8-decode256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_7b_int8.json 2> gpurun_out/bench_7b_int8.err; tail -c 1500 gpurun_out/bench_7b_int8.json; tail -5 gpurun_out/bench_7b_int8.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json

# round 2, call A: full GPU test suite, smoke, the driver's bench command (both arms), baseline timelines
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x --durations=12 ) > gpurun_out/r02a_pytest.log 2>&1; tail -30 gpurun_out/r02a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 6000 gpurun_out/r02a_bench.json; tail -8 gpurun_out/r02a_bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; tail -c 1500 gpurun_out/r02a_bench_ref.json; tail -5 gpurun_out/r02a_bench_ref.err
timeout 250 python scripts/timeline.py llama7b 3 16 > gpurun_out/r02a_timeline_llama7b.txt 2>&1; cat gpurun_out/r02a_timeline_llama7b.txt
timeout 250 python scripts/timeline.py llama7b 3 260 > gpurun_out/r02a_timeline_llama7b_t260.txt 2>&1; cat gpurun_out/r02a_timeline_llama7b_t260.txt

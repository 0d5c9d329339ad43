cd $GRAFT_REPO_ROOT
for v in S0 S1 S2 S0M; do TAG=$v TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_$v.so timeout 120 python scripts/ab_bench.py llama7b 256 2>&1 | tail -1; done
for v in S0 S2; do echo "== timeline $v"; TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_$v.so timeout 200 python scripts/timeline.py llama7b 3 100 2>&1 | sed -n 10,12p; done

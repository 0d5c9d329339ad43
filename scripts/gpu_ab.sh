cd $GRAFT_REPO_ROOT
for v in A B C E F; do TAG=$v TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_$v.so timeout 120 python scripts/ab_bench.py llama7b 256 2>&1 | tail -1; done
for v in A F; do TAG=${v}_oldattn TURBOINFER_B200_LEAN_ATTN=0 TURBOINFER_B200_LIB=$PWD/turboinfer_b200/variants/lib_$v.so timeout 120 python scripts/ab_bench.py llama7b 256 2>&1 | tail -1; done

cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 250 python scripts/timeline.py tinyllama 2 16 2>&1 | tail -14 | head -12
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tinyllama', d['value'], d['whole_step']['us_per_token'])"
timeout 600 python bench.py --workload llama7b-int4-decode256 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('7b', d['value'], d['whole_step']['us_per_token'])"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_baseline_configs_gpu.py -q > $O/s21_pytest.txt 2>&1
tail -2 $O/s21_pytest.txt
for i in 1 2; do timeout 100 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/s21_bench$i.json 2> $O/s21_bench$i.err; python - <<PY
import json
for l in open("$O/s21_bench$i.json"):
    if l.startswith("{"):
        d = json.loads(l); print(round(d["ms_per_step"], 3), round(d["value"]), round(d["roofline"]["gemm_ms_per_step"], 3))
PY
done
timeout 100 python bench.py --precision tf32 --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('tf32', round(d['ms_per_step'],3), round(d['value']))
"

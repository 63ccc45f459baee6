#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/s18_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["ms_per_step"], d["value"], d["e2e"]["value"], json.dumps(d["e2e_loader"])[:900])
PY
tail -3 gpurun_out/s18_bench.err

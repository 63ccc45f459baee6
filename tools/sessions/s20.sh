#!/bin/bash
# last check of the committed state: full GPU suite, smoke, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/s20_pytest.txt 2>&1
tail -2 $O/s20_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/s20_smoke.txt 2>&1
tail -1 $O/s20_smoke.txt
timeout 400 python bench.py > $O/s20_bench.json 2> $O/s20_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/s20_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_loader"]["device_collate"]["value"], d["energy_eval"]["value"], d["cpu_baseline"]["value"], d["clocks"])
PY

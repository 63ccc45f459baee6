#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/m4_bench.json 2> gpurun_out/m4_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/m4_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d["param_sync"]; print(round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), p["identical_on_all_ranks"], p["peer_barrier_timed_out"], p["gradient_exchange"][:30])
PY

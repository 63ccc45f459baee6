#!/bin/bash
# 8-GPU session e: split-barrier peer all-reduce
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/m8e_$tag.json 2> $O/m8e_$tag.err; python - <<PY
import json
for l in open("$O/m8e_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d["param_sync"]; print("$tag", round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), p["identical_on_all_ranks"], p["peer_barrier_timed_out"])
PY
}
run split16 X=1
run split32 GRAPPA_B200_PEER_CTAS=32

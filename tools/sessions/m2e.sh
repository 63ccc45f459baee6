#!/bin/bash
# 2-GPU session e: peer all-reduce with the barriers split out of the data kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_peer_allreduce_gpu.py -x -q > $O/m2e_pytest.txt 2>&1
tail -4 $O/m2e_pytest.txt
run() { tag=$1; shift; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/m2e_$tag.json 2> $O/m2e_$tag.err; python - <<PY
import json
for l in open("$O/m2e_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d["param_sync"]; print("$tag", round(d["ms_per_step"], 3), round(d["value"]), p["identical_on_all_ranks"], p["peer_barrier_timed_out"])
PY
}
run split16 X=1
run mono16 GRAPPA_B200_PEER_SPLIT=0
run split32 GRAPPA_B200_PEER_CTAS=32
run noreduce GRAPPA_B200_SKIP_ALLREDUCE=1

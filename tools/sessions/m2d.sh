#!/bin/bash
# 2-GPU session d: NVLink peer-memory all-reduce
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_peer_allreduce_gpu.py -x -q > $O/m2d_pytest.txt 2>&1
tail -15 $O/m2d_pytest.txt
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/m2d_$tag.json 2> $O/m2d_$tag.err; python - <<PY
import json
for l in open("$O/m2d_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$tag", round(d["ms_per_step"], 3), round(d["value"]), d["param_sync"])
PY
}
run peer16 X=1
run peer8 GRAPPA_B200_PEER_CTAS=8
run peer32 GRAPPA_B200_PEER_CTAS=32
run nccl GRAPPA_B200_PEER_ALLREDUCE=0
run noreduce GRAPPA_B200_SKIP_ALLREDUCE=1
GRAPPA_B200_TRACE=m2d_trace.json timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/step_timeline.py 2>&1 | grep -E "peer_allreduce|optimizer tail|last compute|^step" > $O/m2d_timeline.txt
rm -f $O/m2d_trace.json
cat $O/m2d_timeline.txt | cut -c1-300

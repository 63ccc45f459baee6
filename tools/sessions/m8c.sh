#!/bin/bash
# 8-GPU session c: part-wise writer buckets + NCCL settings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/m8c_$tag.json 2> $O/m8c_$tag.err; python - <<PY
import json
for l in open("$O/m8c_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$tag", round(d["ms_per_step"], 3), round(d["value"]), d["param_sync"]["max_cross_rank_checksum_difference"])
PY
}
run base X=1
run noreduce GRAPPA_B200_SKIP_ALLREDUCE=1
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run nvls NCCL_ALGO=NVLS
run ch16 NCCL_MIN_NCHANNELS=16 NCCL_MAX_NCHANNELS=16
run ch32 NCCL_MIN_NCHANNELS=32 NCCL_MAX_NCHANNELS=32
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep -i -E "NCCL INFO.*(nvls|channels|algo|proto|Connected)" | sed 's/.*NCCL INFO//' | sort | uniq -c | sort -rn | head -25 > $O/m8c_nccl_info.txt
GRAPPA_B200_TRACE=m8c_trace.json timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 tools/step_timeline.py 2>&1 | grep -E "NCCL kernels|optimizer tail|last compute|^step|nccl" > $O/m8c_timeline.txt
rm -f $O/m8c_trace.json

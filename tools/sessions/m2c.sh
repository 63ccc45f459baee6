#!/bin/bash
# 2-GPU session c: where does the data-parallel overhead come from
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/m2c_$tag.json 2> $O/m2c_$tag.err; python - <<PY
import json
for l in open("$O/m2c_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$tag", round(d["ms_per_step"], 3), round(d["value"]), d["param_sync"]["max_cross_rank_checksum_difference"])
PY
}
run base X=1
run noreduce GRAPPA_B200_SKIP_ALLREDUCE=1
run ch2 NCCL_MAX_NCHANNELS=2
run ch4 NCCL_MAX_NCHANNELS=4
run ch8 NCCL_MIN_NCHANNELS=8 NCCL_MAX_NCHANNELS=8
run simple NCCL_PROTO=Simple
run tree NCCL_ALGO=Tree
timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/m2c_1gpu.json 2> $O/m2c_1gpu.err
python -c "
import json
for l in open('$O/m2c_1gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('1gpu', round(d['ms_per_step'],3), round(d['value']))
"

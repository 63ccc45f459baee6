#!/bin/bash
# 8-GPU session: the data-parallel step + energy sweep on all GPUs of one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > $O/m8_bench.json 2> $O/m8_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --no-extras --no-cpu-baseline > $O/m8_bench_4.json 2> $O/m8_bench_4.err
tail -c 400 $O/m8_bench.json; tail -c 300 $O/m8_bench_4.json

#!/bin/bash
# 2-GPU session b: writer buckets exchanged part by part
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/m2b_bench.json 2> $O/m2b_bench.err
GRAPPA_B200_TRACE=m2b_trace.json timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/step_timeline.py > $O/m2b_timeline.txt 2>&1
rm -f $O/m2b_trace.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/m2b_bench_1gpu.json 2> $O/m2b_bench_1gpu.err
timeout 600 python -m pytest tests/test_trainer_gpu.py -q > $O/m2b_pytest.txt 2>&1
tail -c 500 $O/m2b_bench.json; tail -3 $O/m2b_pytest.txt; grep -E "NCCL kernels|optimizer tail|last compute|^step" $O/m2b_timeline.txt

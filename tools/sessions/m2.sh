#!/bin/bash
# 2-GPU session: data-parallel step with / without writer stream priorities; timeline of rank 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $O/m2_bench_$tag.json 2> $O/m2_bench_$tag.err; }
run prio GRAPPA_B200_WRITER_PRIO=1
run noprio GRAPPA_B200_WRITER_PRIO=0
GRAPPA_B200_WRITER_PRIO=1 GRAPPA_B200_TRACE=m2_trace.json timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/step_timeline.py > $O/m2_timeline_prio.txt 2>&1
GRAPPA_B200_WRITER_PRIO=0 GRAPPA_B200_TRACE=m2_trace.json timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/step_timeline.py > $O/m2_timeline_noprio.txt 2>&1
rm -f $O/m2_trace.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/m2_bench_1gpu.json 2> $O/m2_bench_1gpu.err
timeout 600 python bench.py --precision tf32 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/m2_bench_1gpu_tf32.json 2> $O/m2_bench_1gpu_tf32.err
tail -2 $O/m2_bench_prio.json | head -c 400

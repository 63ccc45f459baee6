#!/bin/bash
# 8-GPU session b: timeline of rank 0's step, NCCL algorithm report
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
GRAPPA_B200_TRACE=m8_trace.json timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/step_timeline.py > $O/m8_timeline.txt 2>&1
python - <<'PY' > gpurun_out/m8_trace_nccl.txt 2>&1
import json, glob
for f in glob.glob('gpurun_out/m8_trace.json') + glob.glob('m8_trace.json'):
    ev = [e for e in json.load(open(f))['traceEvents'] if e.get('cat') == 'kernel']
    ev.sort(key=lambda e: e['ts'])
    t0 = ev[0]['ts']
    # last step only: take the last 600 kernels
    for e in ev[-620:]:
        n = e['name']
        if 'nccl' in n or 'adam' in n or 'sumsq' in n or 'molwise' in n or 'energy' in n:
            print(f"{e['ts']-t0:10.1f} {e['dur']:8.1f} stream {e['args'].get('stream')} {n[:60]}")
    break
PY
rm -f $O/m8_trace.json m8_trace.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/m8_bench_dbg.json 2> $O/m8_bench_dbg.err
grep -i -E "nvls|algo|channel" $O/m8_bench_dbg.err | sort | uniq -c | sort -rn | head -40 > $O/m8_nccl_info.txt
rm -f $O/m8_bench_dbg.err
head -30 $O/m8_timeline.txt | tail -22

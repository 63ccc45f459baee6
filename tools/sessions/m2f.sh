#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_peer_allreduce_gpu.py tests/test_trainer_gpu.py -x -q > $O/m2f_pytest.txt 2>&1
tail -3 $O/m2f_pytest.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu-baseline > $O/m2f_bench.json 2> $O/m2f_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/m2f_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["param_sync"], d["e2e_loader"]["value"], d["e2e_loader"]["device_collate"]["value"], d["energy_eval"]["value"])
PY
tail -3 $O/m2f_bench.err | cut -c1-300

#!/bin/bash
# device-collate tests, full GPU suite, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_dataset_gpu.py -x -q > $O/s11_dataset.txt 2>&1
tail -15 $O/s11_dataset.txt
timeout 2400 python -m pytest tests -q -m gpu > $O/s11_pytest.txt 2>&1
tail -8 $O/s11_pytest.txt
timeout 900 python bench.py > $O/s11_bench.json 2> $O/s11_bench.err
tail -c 600 $O/s11_bench.json

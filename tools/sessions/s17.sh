#!/bin/bash
# final 1-GPU session of round 2: full GPU suite, smoke, default bench (with extras), reference arm, launch list of the step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/s17_pytest.txt 2>&1
tail -3 $O/s17_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/s17_smoke.txt 2>&1
tail -3 $O/s17_smoke.txt
timeout 600 python bench.py > $O/s17_bench.json 2> $O/s17_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s17_bench_ref.json 2> $O/s17_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/s17_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-cuda-graph > $O/s17_ncu_bench.log 2>&1
tail -c 300 $O/s17_bench.json; tail -c 300 $O/s17_bench_ref.json

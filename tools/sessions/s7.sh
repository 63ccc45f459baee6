#!/bin/bash
# GPU session 7: bf16x3 CTA pairs from K = 128
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bf16x3 or config1 or dropin or trainer" > $O/s7_pytest.log 2>&1
echo "rc=$?" >> $O/s7_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s7_gemm_x3.txt 2>&1
GRAPPA_B200_GEMM_PAIR_MINK_X3=1024 timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s7_gemm_x3_nopair.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s7_bench.json 2> $O/s7_bench.err
GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/step_timeline.py > $O/s7_timeline.txt 2>&1
GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/step_timeline.py --serial > $O/s7_timeline_serial.txt 2>&1
rm -f $O/step_trace.json
tail -4 $O/s7_pytest.log

#!/bin/bash
# GPU session 10: pair-mode forwarder for the converted-stage barrier
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "gemm or bf16x3 or config1 or config2 or dropin or trainer or model" > $O/s10_pytest.log 2>&1
echo "rc=$?" >> $O/s10_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s10_gemm_x3.txt 2>&1
GRAPPA_B200_GEMM_PAIR_MINK_X3=1024 timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s10_gemm_x3_nopair.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s10_bench.json 2> $O/s10_bench.err
GRAPPA_B200_GEMM_PAIR_MINK_X3=1024 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s10_bench_nopair.json 2> $O/s10_bench_nopair.err
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s10_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s10_ncu_gemm_nn.log 2>&1
tail -4 $O/s10_pytest.log

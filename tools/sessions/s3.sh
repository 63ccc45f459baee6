#!/bin/bash
# GPU session 3: converter fence fix; ncu of the packed-pair energy kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bf16x3 or dropin or trainer or config1" > $O/s3_pytest.log 2>&1
echo "rc=$?" >> $O/s3_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s3_gemm_x3.txt 2>&1
timeout 600 python bench.py --precision bf16x3 --steps 20 --warmup 5 --no-cpu-baseline > $O/s3_bench_bf16x3.json 2> $O/s3_bench_bf16x3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_pairs -s 1 -c 1 -o $O/s3_energy_pairs python tools/energy_one.py 1000 100 5 > $O/s3_ncu_energy.log 2>&1
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s3_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s3_ncu_gemm_nn.log 2>&1
tail -4 $O/s3_pytest.log

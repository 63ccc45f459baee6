#!/bin/bash
# GPU session 4: integer bf16 split in the GEMM converter; energy pairs kernel v2 (staged records, split round barrier)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bf16x3 or energy or config1 or config3 or dropin" > $O/s4_pytest.log 2>&1
echo "rc=$?" >> $O/s4_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s4_gemm_x3.txt 2>&1
timeout 300 python tools/energy_bench.py --full > $O/s4_energy.txt 2>&1
GRAPPA_B200_ENERGY_MINB=2 timeout 300 python tools/energy_bench.py > $O/s4_energy_minb2.txt 2>&1
timeout 600 python bench.py --precision bf16x3 --steps 20 --warmup 5 --no-cpu-baseline > $O/s4_bench_bf16x3.json 2> $O/s4_bench_bf16x3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_pairs -s 1 -c 1 -o $O/s4_energy_pairs python tools/energy_one.py 1000 100 5 > $O/s4_ncu_energy.log 2>&1
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s4_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s4_ncu_gemm_nn.log 2>&1
tail -4 $O/s4_pytest.log

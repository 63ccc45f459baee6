#!/bin/bash
# GPU session 5: Veltkamp packed split, backoff waits; new bench.py end to end
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bf16x3 or energy or config1 or config3 or dropin or trainer" > $O/s5_pytest.log 2>&1
echo "rc=$?" >> $O/s5_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s5_gemm_x3.txt 2>&1
timeout 300 python tools/energy_bench.py --full > $O/s5_energy.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s5_bench.json 2> $O/s5_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/s5_bench_ref.json 2> $O/s5_bench_ref.err
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s5_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s5_ncu_gemm_nn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_pairs -s 1 -c 1 -o $O/s5_energy_pairs python tools/energy_one.py 1000 100 5 > $O/s5_ncu_energy.log 2>&1
tail -4 $O/s5_pytest.log

#!/bin/bash
# GPU session 2: shared-space converter, packed-pair energy kernel, ncu captures of both
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/s2_pytest.log 2>&1
echo "rc=$?" >> $O/s2_pytest.log
timeout 300 python tools/gemm_bench.py --precision tf32 > $O/s2_gemm_tf32.txt 2>&1
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s2_gemm_x3.txt 2>&1
timeout 300 python tools/energy_bench.py --full > $O/s2_energy.txt 2>&1
GRAPPA_B200_ENERGY_MINB=2 timeout 300 python tools/energy_bench.py > $O/s2_energy_minb2.txt 2>&1
for prec in tf32 bf16x3; do
  timeout 600 python bench.py --precision "$prec" --steps 20 --warmup 5 --no-cpu-baseline > $O/s2_bench_$prec.json 2> $O/s2_bench_$prec.err
done
# ncu: one bf16x3 GEMM launch (K-major and MN-major operands) and the packed-pair energy kernel, full sets with source
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s2_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s2_ncu_gemm_nn.log 2>&1
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s2_gemm_x3_tt python tools/gemm_one.py 1536 512 14848 1 1 > $O/s2_ncu_gemm_tt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_pairs -s 3 -c 1 -o $O/s2_energy_pairs python tools/energy_one.py 1000 100 5 > $O/s2_ncu_energy.log 2>&1
tail -4 $O/s2_pytest.log

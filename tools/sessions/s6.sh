#!/bin/bash
# GPU session 6: pipe micro-benchmark; Veltkamp split (contraction-proof); energy backoff; writer priorities
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/microbench/pipes.cu > $O/s6_pipes_build.log 2>&1 && /tmp/pipes > $O/s6_pipes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "bf16x3 or energy or config1 or config3 or dropin or trainer" > $O/s6_pytest.log 2>&1
echo "rc=$?" >> $O/s6_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s6_gemm_x3.txt 2>&1
timeout 300 python tools/energy_bench.py > $O/s6_energy.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s6_bench.json 2> $O/s6_bench.err
GRAPPA_B200_WRITER_PRIO=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s6_bench_noprio.json 2> $O/s6_bench_noprio.err
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s6_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s6_ncu_gemm_nn.log 2>&1
tail -4 $O/s6_pytest.log

#!/bin/bash
# GPU session 15: bf16x3 converter split A/B (Veltkamp vs cvt.rn.bf16x2) x warp roles
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for v in v1 v2 v3 v5; do
  export GRAPPA_B200_LIB=$PWD/tools/_trace/$v.so
  timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s15_gemm_x3_$v.txt 2>&1
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s15_bench_$v.json 2> $O/s15_bench_$v.err
done
export GRAPPA_B200_LIB=$PWD/tools/_trace/v2.so
timeout 900 python -m pytest tests -m gpu -q -k "gemm or bf16x3 or config1 or config2" > $O/s15_pytest_v2.log 2>&1
export GRAPPA_B200_LIB=$PWD/tools/_trace/v2t.so
GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/gemm_trace.py 14848 1536 512 > $O/s15_trace_nn_v2t.txt 2>&1
unset GRAPPA_B200_LIB
for v in v1 v2 v3 v5; do tail -1 $O/s15_gemm_x3_$v.txt; tail -c 200 $O/s15_bench_$v.json | head -c 10; done
tail -3 $O/s15_pytest_v2.log

#!/bin/bash
# GPU session 14: bf16x3 warp-role A/B (8 epilogue warps / converter grouping), swizzled epilogue patch, 6-stage rings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "gemm or bf16x3 or config1 or config2 or dropin or trainer or model" > $O/s14_pytest.log 2>&1
echo "rc=$?" >> $O/s14_pytest.log
for v in libgrappa_b200_trace lib_c2g3 lib_e4; do
  export GRAPPA_B200_LIB=$PWD/tools/_trace/$v.so
  timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s14_gemm_x3_$v.txt 2>&1
  GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/gemm_trace.py 14848 1536 512 > $O/s14_trace_nn_$v.txt 2>&1
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s14_bench_$v.json 2> $O/s14_bench_$v.err
done
unset GRAPPA_B200_LIB
timeout 300 python tools/gemm_bench.py --precision tf32 > $O/s14_gemm_tf32.txt 2>&1
timeout 900 python bench.py --precision tf32 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s14_bench_tf32.json 2> $O/s14_bench_tf32.err
tail -4 $O/s14_pytest.log; for v in libgrappa_b200_trace lib_c2g3 lib_e4; do tail -1 $O/s14_gemm_x3_$v.txt; done

#!/bin/bash
# GPU session 13: cluster arrivals without the .release.cluster fence (converters arrive on the leader directly)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "gemm or bf16x3 or config1 or config2 or dropin or trainer or model" > $O/s13_pytest.log 2>&1
echo "rc=$?" >> $O/s13_pytest.log
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/s13_gemm_x3.txt 2>&1
timeout 300 python tools/gemm_bench.py --precision tf32 > $O/s13_gemm_tf32.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s13_bench.json 2> $O/s13_bench.err
timeout 900 python bench.py --precision tf32 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/s13_bench_tf32.json 2> $O/s13_bench_tf32.err
export GRAPPA_B200_LIB=$PWD/tools/_trace/libgrappa_b200_trace.so
GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/gemm_trace.py 14848 1536 512 > $O/s13_trace_bf16x3_nn.txt 2>&1
GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/gemm_trace.py 14848 512 512 0 1 > $O/s13_trace_bf16x3_nt.txt 2>&1
GRAPPA_B200_PREC=bf16x3 timeout 300 python tools/gemm_trace.py 512 512 14848 1 1 > $O/s13_trace_bf16x3_tt.txt 2>&1
unset GRAPPA_B200_LIB
tail -4 $O/s13_pytest.log; tail -c 300 $O/s13_bench.json

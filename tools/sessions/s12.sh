#!/bin/bash
# pipeline timelines of single GEMM launches (instrumented debug library)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
export GRAPPA_B200_LIB=$PWD/tools/_trace/libgrappa_b200_trace.so
for prec in bf16x3 tf32; do
  GRAPPA_B200_PREC=$prec timeout 300 python tools/gemm_trace.py 14848 1536 512 > $O/s12_trace_${prec}_nn.txt 2>&1
  GRAPPA_B200_PREC=$prec timeout 300 python tools/gemm_trace.py 14848 512 512 0 1 > $O/s12_trace_${prec}_nt.txt 2>&1
done
GRAPPA_B200_PREC=bf16x3 GRAPPA_B200_GEMM_PAIR=0 timeout 300 python tools/gemm_trace.py 14848 1536 512 > $O/s12_trace_bf16x3_nn_single.txt 2>&1
unset GRAPPA_B200_LIB
timeout 600 python -m pytest tests/test_model_gpu.py -x -q > $O/s12_pytest_model.txt 2>&1
tail -3 $O/s12_pytest_model.txt
tail -5 $O/s12_trace_bf16x3_nn.txt

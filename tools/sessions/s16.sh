#!/bin/bash
# GPU session 16: ncu of the bf16x3 GEMM (cvt split, 8 epilogue + 3x2 converter warps), full GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
GRAPPA_B200_PREC=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/s16_gemm_x3_nn python tools/gemm_one.py 14848 1536 512 > $O/s16_ncu_gemm_nn.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/s16_pytest.log 2>&1
tail -3 $O/s16_pytest.log

#!/bin/bash
# GPU session 1 (round 2): state after the advisor fixes + the bf16x3 GEMM; error budget; step time per precision
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/s1_smi.txt 2>&1
# (1) everything that does not touch the new kernel
timeout 900 python -m pytest tests -m gpu -q -x -k "not bf16x3" --deselect "tests/test_baseline_configs_gpu.py::test_config1_training_batch_loss_and_backward_match_reference_fixture[bench]" > $O/s1_pytest_base.log 2>&1
echo "base rc=$?" >> $O/s1_pytest_base.log
# (2) the new kernel's unit tests (own process: a trap would poison the CUDA context)
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "bf16x3" > $O/s1_pytest_x3.log 2>&1
echo "x3 rc=$?" >> $O/s1_pytest_x3.log
# (3) full-size backward at the bench precision
timeout 600 python -m pytest "tests/test_baseline_configs_gpu.py::test_config1_training_batch_loss_and_backward_match_reference_fixture" -m gpu -q -s > $O/s1_pytest_full.log 2>&1
echo "full rc=$?" >> $O/s1_pytest_full.log
# (4) error budget
timeout 900 python tools/error_budget.py --out $O/s1_error_budget.md > $O/s1_error_budget.log 2>&1
# (5) per-shape GEMM timing, both arithmetics
timeout 300 python tools/gemm_bench.py --precision tf32 --check > $O/s1_gemm_tf32.txt 2>&1
timeout 300 python tools/gemm_bench.py --precision bf16x3 --check > $O/s1_gemm_x3.txt 2>&1
# (6) step time per arithmetic
for prec in tf32 bf16x3 "fwd=bf16x3,dgrad=bf16x3,wgrad=tf32" "fwd=bf16x3,dgrad=tf32,wgrad=tf32"; do
  tag=$(echo $prec | tr '=,' '__')
  timeout 600 python bench.py --precision "$prec" --steps 20 --warmup 5 --no-cpu-baseline > $O/s1_bench_$tag.json 2> $O/s1_bench_$tag.err
done
# (7) timelines
GRAPPA_B200_PREC=bf16x3 GRAPPA_B200_TRACE=s1_trace_x3.json timeout 300 python tools/step_timeline.py > $O/s1_timeline_x3.txt 2>&1
GRAPPA_B200_PREC=bf16x3 GRAPPA_B200_TRACE=s1_trace_x3_serial.json timeout 300 python tools/step_timeline.py --serial > $O/s1_timeline_x3_serial.txt 2>&1
rm -f $O/s1_trace_x3.json $O/s1_trace_x3_serial.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/s1_smoke.log 2>&1
tail -3 $O/s1_smoke.log $O/s1_pytest_base.log $O/s1_pytest_x3.log $O/s1_pytest_full.log

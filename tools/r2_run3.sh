#!/bin/bash
# round 2: kernels + rep + fused first, then whole suite, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_rep_gpu.py tests/test_fused_step_gpu.py tests/test_engine_state_gpu.py -m gpu -q --timeout=600 --timeout-method=thread -s > gpurun_out/r2_tests3a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests3a.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests3a.log | tail -100
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 --timeout-method=thread -s --deselect tests/test_kernels_gpu.py --deselect tests/test_rep_gpu.py --deselect tests/test_fused_step_gpu.py --deselect tests/test_engine_state_gpu.py > gpurun_out/r2_tests3b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests3b.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests3b.log | tail -70
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench3.log 2>&1
tail -2 gpurun_out/r2_bench3.log | cut -c1-3000

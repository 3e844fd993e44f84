#!/bin/bash
# round 2, call 2: representation-path kernels + fused step: new tests first (fail fast output), then the whole suite, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rep_gpu.py tests/test_fused_step_gpu.py tests/test_engine_state_gpu.py -m gpu -q --timeout=600 --timeout-method=thread -s > gpurun_out/r2_tests2a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests2a.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests2a.log | tail -120
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 --timeout-method=thread -s --deselect tests/test_rep_gpu.py --deselect tests/test_fused_step_gpu.py --deselect tests/test_engine_state_gpu.py > gpurun_out/r2_tests2b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests2b.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests2b.log | tail -60
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench2.log 2>&1
tail -2 gpurun_out/r2_bench2.log | cut -c1-1200

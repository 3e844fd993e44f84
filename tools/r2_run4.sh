#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_full_width_gpu.py tests/test_model_gpu.py tests/test_fused_step_gpu.py -m gpu -q --timeout=600 --timeout-method=thread -x > gpurun_out/r2_tests4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests4.log | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-ddim > gpurun_out/r2_bench4.log 2>&1
tail -2 gpurun_out/r2_bench4.log | cut -c1-2600

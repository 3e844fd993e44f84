#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rep_gpu.py tests/test_kernels_gpu.py tests/test_fused_step_gpu.py tests/test_full_width_gpu.py tests/test_model_gpu.py tests/test_samplers_gpu.py tests/test_api_variants_gpu.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/r2_tests14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests14.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests14.log | tail -40
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-ref > gpurun_out/r2_bench14.log 2>&1
tail -2 gpurun_out/r2_bench14.log | cut -c1-330
grep -o '"ddim": {.*' gpurun_out/r2_bench14.log | cut -c1-600

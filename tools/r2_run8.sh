#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_api_variants_gpu.py tests/test_evaluation_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/r2_tests8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests8.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests8.log | tail -60

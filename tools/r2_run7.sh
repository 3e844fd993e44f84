#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_samplers_gpu.py tests/test_full_width_gpu.py tests/test_train_parity_gpu.py -m gpu -q --timeout=600 --timeout-method=thread -s > gpurun_out/r2_tests7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests7.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests7.log | grep -v "^input_blocks\|^output_blocks\|^middle" | tail -60
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench7.log 2>&1
tail -2 gpurun_out/r2_bench7.log | cut -c1-400
grep -o '"ddim": {.*' gpurun_out/r2_bench7.log | cut -c1-700

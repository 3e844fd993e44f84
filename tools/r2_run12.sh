#!/bin/bash
mkdir -p gpurun_out
CDAE_T_CFG=3 timeout 300 python tools/gpu_igemm_bench.py fwd stats > gpurun_out/r2_igemm_bench_t333.log 2>&1
head -6 gpurun_out/r2_igemm_bench_t333.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_full_width_gpu.py -m gpu -q --timeout=300 --timeout-method=thread -x 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench12.log 2>&1
tail -2 gpurun_out/r2_bench12.log | cut -c1-330
grep -o '"roofline": {.*"hbm_kernels"' gpurun_out/r2_bench12.log | cut -c1-900
grep -o '"ddim": {.*' gpurun_out/r2_bench12.log | cut -c1-600

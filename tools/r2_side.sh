#!/bin/bash
mkdir -p gpurun_out
CDAE_T_CFG=5 timeout 300 python tools/gpu_igemm_bench.py fwd stats > gpurun_out/r2_igemm_bench_t253_stats.log 2>&1
timeout 600 python -m pytest tests/test_fused_step_gpu.py tests/test_full_width_gpu.py tests/test_model_gpu.py -m gpu -q --timeout=300 --timeout-method=thread -x 2>&1 | tail -3
CDAE_WGRAD_SIDE_STREAM=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-gpu-ref --no-cfg1 --no-ddim > gpurun_out/r2_bench_side0.log 2>&1
tail -2 gpurun_out/r2_bench_side0.log | cut -c1-260
CDAE_WGRAD_SIDE_STREAM=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-gpu-ref --no-cfg1 --no-ddim > gpurun_out/r2_bench_side1.log 2>&1
tail -2 gpurun_out/r2_bench_side1.log | cut -c1-260

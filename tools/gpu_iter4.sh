#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x --timeout=300 --timeout-method=thread 2>&1 | tail -12 | tee gpurun_out/iter_tests.log
timeout 300 python tools/gpu_igemm_bench.py fwd 2>&1 | tee gpurun_out/igemm_bench.log
CDAE_NO_HALO=1 timeout 300 python tools/gpu_igemm_bench.py fwd 2>&1 | tee gpurun_out/igemm_bench_nohalo.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | grep '^{' | tee gpurun_out/bench_iter.log
timeout 600 python tools/gpu_hostprof.py 2>&1 | head -40 | cut -c1-200 | tee gpurun_out/hostprof.log

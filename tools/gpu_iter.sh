#!/bin/bash
# quick iteration: igemm/wgrad parity tests + micro-benchmark
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout=120 --timeout-method=thread -k "igemm or wgrad" 2>&1 | tail -15 | tee gpurun_out/iter_tests.log
timeout 300 python tools/gpu_igemm_bench.py $1 2>&1 | tee gpurun_out/igemm_bench.log

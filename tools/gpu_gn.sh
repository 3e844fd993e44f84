#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm" --timeout=300 2>&1 | tail -3 | tee gpurun_out/gn_tests.log
timeout 300 python tools/gpu_gn_bench.py 2>&1 | tee gpurun_out/gn_bench.log

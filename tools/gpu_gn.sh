#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm" --timeout=300 2>&1 | tail -5 | tee gpurun_out/gn_tests.log
CDAE_GN_DEBUG=1 timeout 300 python tools/gpu_gn_bench.py 2>&1 | sort -u | tee gpurun_out/gn_bench.log

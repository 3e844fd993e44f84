#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" --timeout=300 2>&1 | tail -5 | tee gpurun_out/attn_tests.log
timeout 300 python tools/gpu_attn_bench.py 2>&1 | tee gpurun_out/attn_bench.log

#!/bin/bash
mkdir -p gpurun_out
CDAE_GN_DEBUG=1 CDAE_GN_MIN_ROW=32 timeout 300 python tools/gpu_gn_bench.py 2>&1 | sort -u | grep -v "^gn bwd" | tee gpurun_out/gn_bench_row32.log
timeout 300 python tools/gpu_gn_bench.py 2>&1 | tee gpurun_out/gn_bench.log

#!/bin/bash
# ncu --set full of the dominant tensor-core kernels on their heaviest cfg2 shapes (first launches of the micro-benchmark)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm3_kernel -c 2 -o gpurun_out/prof_igemm3_128c -f python tools/gpu_igemm_bench.py fwd > gpurun_out/ncu_igemm3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad3_kernel -c 1 -o gpurun_out/prof_wgrad3 -f python tools/gpu_igemm_bench.py wg > gpurun_out/ncu_wgrad3.log 2>&1
ls -la gpurun_out | grep prof

#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm3t_kernel -s 5 -c 1 -o gpurun_out/r2_ncu_igemm3t -f python tools/gpu_igemm_bench.py fwd first > gpurun_out/r2_ncu_t.log 2>&1
tail -2 gpurun_out/r2_ncu_t.log

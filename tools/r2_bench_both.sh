#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/r2_bench_full.log 2>&1
tail -5 gpurun_out/r2_bench_full.log | cut -c1-7000
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/r2_bench_reference.log 2>&1
tail -5 gpurun_out/r2_bench_reference.log | cut -c1-2500
CDAE_T_CFG=5 timeout 300 python tools/gpu_igemm_bench.py fwd stats 2>&1 | head -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm3t_kernel -s 5 -c 1 -o gpurun_out/r2_ncu_igemm3t -f python tools/gpu_igemm_bench.py fwd first > gpurun_out/r2_ncu_t.log 2>&1
tail -1 gpurun_out/r2_ncu_t.log

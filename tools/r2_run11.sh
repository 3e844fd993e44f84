#!/bin/bash
mkdir -p gpurun_out
CDAE_IGEMM3T_ALL=1 timeout 300 python tools/gpu_igemm_bench.py fwd > gpurun_out/r2_igemm_bench_tall.log 2>&1
head -8 gpurun_out/r2_igemm_bench_tall.log
CDAE_IGEMM3T_ALL=1 timeout 300 python tools/gpu_igemm_bench.py fwd stats > gpurun_out/r2_igemm_bench_tall_stats.log 2>&1
head -6 gpurun_out/r2_igemm_bench_tall_stats.log

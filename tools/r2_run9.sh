#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout=300 --timeout-method=thread -x > gpurun_out/r2_tests9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests9.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests9.log | tail -40
timeout 300 python tools/r2_gnb_bench.py > gpurun_out/r2_gnb_bench_t.log 2>&1
cat gpurun_out/r2_gnb_bench_t.log
timeout 300 python tools/gpu_igemm_bench.py fwd > gpurun_out/r2_igemm_bench_t.log 2>&1
head -8 gpurun_out/r2_igemm_bench_t.log
timeout 300 python tools/gpu_igemm_bench.py fwd stats > gpurun_out/r2_igemm_bench_t_stats.log 2>&1
head -6 gpurun_out/r2_igemm_bench_t_stats.log

#!/bin/bash
# round-1 iteration 5: GroupNorm statistics from the conv epilogue + streaming forward GroupNorm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread -x 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -25 | tee gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench.log
timeout 300 python tools/gpu_gn_bench.py 2>&1 | tee gpurun_out/gn_bench.log
timeout 200 python tools/gpu_igemm_bench.py fwd 2>&1 | tee gpurun_out/igemm_bench.log
timeout 200 python tools/gpu_igemm_bench.py fwd stats 2>&1 | tee gpurun_out/igemm_bench_stats.log
CDAE_FUSED_GN_STATS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-ddim --no-cpu 2>&1 | grep '^{' | tee gpurun_out/bench_nofuse.log

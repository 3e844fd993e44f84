#!/bin/bash
# state check: tests + smoke + bench + per-op profile + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -x 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -15 | tee gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench.log
timeout 300 python tools/gpu_opprof.py 2>&1 | tail -50 | tee gpurun_out/opprof.log
timeout 300 python tools/gpu_igemm_bench.py fwd 2>&1 | tee gpurun_out/igemm_bench.log
timeout 300 python tools/gpu_gn_bench.py 2>&1 | tee gpurun_out/gn_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-ddim --no-cpu > gpurun_out/ncu_bench_stdout.log 2>&1
tail -2 gpurun_out/ncu_bench_stdout.log

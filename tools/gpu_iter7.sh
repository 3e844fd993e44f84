#!/bin/bash
# round-1 iteration 7: streaming two-pass GroupNorm backward (A/B against the resident kernel), evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread -x 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -25 | tee gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench.log
CDAE_GN_BWD_STREAM=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-ddim --no-cpu 2>&1 | grep '^{' | tee gpurun_out/bench_bwd_resident.log
timeout 300 python tools/gpu_gn_bench.py 2>&1 | tee gpurun_out/gn_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-ddim --no-cpu > gpurun_out/ncu_bench_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_reduce_kernel -s 5 -c 1 -o gpurun_out/ncu_gn_bwd_reduce -f \
    python tools/gpu_gn_bench.py first > gpurun_out/ncu_gn_bwd_reduce_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_apply_kernel -s 5 -c 1 -o gpurun_out/ncu_gn_bwd_apply -f \
    python tools/gpu_gn_bench.py first > gpurun_out/ncu_gn_bwd_apply_stdout.log 2>&1
ls -la gpurun_out | tail -8

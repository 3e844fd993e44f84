#!/bin/bash
# round-1 iteration 6: pipelined gn_apply, data path, checkpoint/resume; ncu evidence for the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread -x 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -25 | tee gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench.log
timeout 300 python tools/gpu_gn_bench.py 2>&1 | tee gpurun_out/gn_bench.log
timeout 200 python tools/gpu_data_bench.py 2>&1 | grep '^{' | tee gpurun_out/data_bench.log
# launch list of ~2 steady-state training steps (share of the step per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-ddim --no-cpu > gpurun_out/ncu_bench_stdout.log 2>&1
tail -2 gpurun_out/ncu_bench_stdout.log | cut -c1-300
# full captures: streaming GroupNorm forward, conv with the statistics epilogue, dataset gather
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_apply_fwd_kernel -s 5 -c 2 -o gpurun_out/ncu_gn_apply -f \
    python tools/gpu_gn_bench.py first > gpurun_out/ncu_gn_apply_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm3_kernel -s 6 -c 2 -o gpurun_out/ncu_igemm3_stats -f \
    python tools/gpu_igemm_bench.py fwd stats first > gpurun_out/ncu_igemm3_stats_stdout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_images_kernel -s 10 -c 1 -o gpurun_out/ncu_gather -f \
    python tools/gpu_data_bench.py first > gpurun_out/ncu_gather_stdout.log 2>&1
ls -la gpurun_out | tail -20

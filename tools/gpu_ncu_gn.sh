#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_fwd_pipe -c 1 -o gpurun_out/prof_gn_fwd -f python tools/gpu_gn_bench.py > gpurun_out/ncu_gn_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_pipe -c 1 -o gpurun_out/prof_gn_bwd -f python tools/gpu_gn_bench.py > gpurun_out/ncu_gn_bwd.log 2>&1
ls -la gpurun_out

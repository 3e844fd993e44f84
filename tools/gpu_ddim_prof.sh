#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_ddim_prof.py 128 2>&1 | tail -1 | tee gpurun_out/ddim_prof.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ddim.csv python tools/gpu_ddim_prof.py 128 > gpurun_out/ncu_ddim_stdout.log 2>&1
tail -1 gpurun_out/ncu_ddim_stdout.log

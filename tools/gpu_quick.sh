#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_hostprof.py 2>&1 | tail -75 | tee gpurun_out/hostprof.log

#!/bin/bash
# descriptor experiment sweep + per-op profile of one training step
mkdir -p gpurun_out
for mode in 0 1; do for G in 8 10 16; do for r0 in 0 1 3 8 11; do for v in 0 1; do
  timeout 20 tools/exp_desc.bin $mode $r0 $G $v
done; done; done; done 2>&1 | tee gpurun_out/exp_desc.log
timeout 600 python tools/gpu_opprof.py 2>&1 | tail -50 | tee gpurun_out/opprof.log

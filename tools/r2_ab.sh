#!/bin/bash
# A/B of one environment switch on the training step: tools/r2_ab.sh VAR   (two interleaved repetitions)
mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export $1=1; else unset $1; fi
  timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-gpu-ref --no-cfg1 --no-ddim > gpurun_out/ab_$v.log 2>&1
  echo "$1=$v $(tail -1 gpurun_out/ab_$v.log | grep -o '"ms_per_step": [0-9.]*' | head -1)"
done
done

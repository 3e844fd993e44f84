#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm3_kernel -s 17 -c 1 -o gpurun_out/r2_ncu_igemm3_gnb -f python tools/r2_gnb_bench.py first > gpurun_out/r2_ncu_gnb.log 2>&1
tail -2 gpurun_out/r2_ncu_gnb.log

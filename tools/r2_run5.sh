#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/r2_gnb_bench.py > gpurun_out/r2_gnb_bench.log 2>&1
cat gpurun_out/r2_gnb_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm3_kernel -s 7 -c 1 -o gpurun_out/r2_ncu_igemm3_gnb -f python tools/r2_gnb_bench.py first > gpurun_out/r2_ncu_gnb.log 2>&1
tail -3 gpurun_out/r2_ncu_gnb.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_apply -s 2 -c 1 -o gpurun_out/r2_ncu_gn_bwd_apply -f python tools/r2_gnb_bench.py first > gpurun_out/r2_ncu_gnapply.log 2>&1
tail -3 gpurun_out/r2_ncu_gnapply.log

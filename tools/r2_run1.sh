#!/bin/bash
# round 2, call 1: whole GPU suite (new full-width parity tests print their measured errors) + a baseline bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 --timeout-method=thread -s --durations=20 > gpurun_out/r2_tests1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests1.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests1.log | tail -150
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench1.log 2>&1
tail -2 gpurun_out/r2_bench1.log

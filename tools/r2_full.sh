#!/bin/bash
# round 2 full validation on one B200: whole GPU suite, smoke, both bench arms with the default arguments
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 --timeout-method=thread -s --durations=12 > gpurun_out/r2_tests_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_full.log
grep -vE "^\s*$|UserWarning|_warn|return float" gpurun_out/r2_tests_full.log | grep -v "^input_blocks\|^output_blocks\|^middle" | tail -45
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench_full.log 2>&1
tail -6 gpurun_out/r2_bench_full.log | cut -c1-6000
( time timeout 1200 python bench.py --impl reference ) > gpurun_out/r2_bench_reference.log 2>&1
tail -6 gpurun_out/r2_bench_reference.log | cut -c1-2500

#!/bin/bash
# the whole GPU suite (no bench): quick regression gate between kernel changes
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 --timeout-method=thread -x > gpurun_out/r2_tests_quick.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_quick.log
grep -E "passed|failed|rc=|^E |Error" gpurun_out/r2_tests_quick.log | tail -12

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -x 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -15 | tee gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench.log

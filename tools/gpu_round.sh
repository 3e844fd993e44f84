#!/bin/bash
# tests + smoke + bench on one B200; everything bounded by timeouts
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -x -s 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -40 | tee gpurun_out/tests.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log

#!/bin/bash
# round-1 iteration 8: shipped-configuration parity tests, image_train recipe, bench with the GroupNorm roofline entries
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread 2>&1 | grep -vE "^\s*$|UserWarning|_warn|return float" | tail -60 | tee gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench.log

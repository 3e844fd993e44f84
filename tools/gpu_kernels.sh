#!/bin/bash
# first-contact GPU run: kernel-level parity tests, each bounded so a hung kernel cannot wedge the box
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout=180 --timeout-method=thread 2>&1 | tail -120 | tee gpurun_out/kernels.log

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 --timeout-method=thread -x 2>&1 | tail -80 | tee gpurun_out/model.log

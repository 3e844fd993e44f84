#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=800 --timeout-method=thread > gpurun_out/r2_tests_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_2gpu.log
grep -E "DISTGRAD|passed|failed|rc=" gpurun_out/r2_tests_2gpu.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-ddim --no-cpu > gpurun_out/r2_bench_2gpu_bf16wire.log 2>&1
tail -2 gpurun_out/r2_bench_2gpu_bf16wire.log | cut -c1-330
CDAE_GRAD_WIRE=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-ddim --no-cpu > gpurun_out/r2_bench_2gpu_fp32wire.log 2>&1
tail -2 gpurun_out/r2_bench_2gpu_fp32wire.log | cut -c1-330

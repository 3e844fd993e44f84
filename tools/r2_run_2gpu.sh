#!/bin/bash
# 2 x B200: NCCL gradient parity (fp32 and bf16 wire, two-graph step with the early all-reduce) + the bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=800 --timeout-method=thread > gpurun_out/r2_tests_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_2gpu.log
grep -E "DISTGRAD|passed|failed|rc=|Error" gpurun_out/r2_tests_2gpu.log | cut -c1-400
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-ddim --no-cpu > gpurun_out/$2 2>&1; grep '^{' gpurun_out/$2 | tail -1 | cut -c1-230; }
echo "bf16 wire, overlapped"; run 29511 r2_bench_2gpu_bf16wire.log
echo "bf16 wire, not overlapped"; CDAE_OVERLAP_ALLREDUCE=0 run 29512 r2_bench_2gpu_bf16wire_nooverlap.log
echo "fp32 wire, overlapped"; CDAE_GRAD_WIRE=fp32 run 29513 r2_bench_2gpu_fp32wire.log
echo "one GPU, same box"; timeout 600 python bench.py --steps 20 --warmup 5 --no-ddim --no-cpu --no-gpu-ref --no-cfg1 > gpurun_out/r2_bench_1gpu_samebox.log 2>&1; grep '^{' gpurun_out/r2_bench_1gpu_samebox.log | tail -1 | cut -c1-230

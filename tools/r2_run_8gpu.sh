#!/bin/bash
# N = 8 (or whatever the box has): the bench line through torchrun, overlapped and not
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-ddim --no-cpu > gpurun_out/$2 2>&1; grep '^{' gpurun_out/$2 | tail -1 | cut -c1-220; }
echo "bf16 wire, overlapped"; run 29521 r2_bench_${N}gpu.log
echo "bf16 wire, not overlapped"; CDAE_OVERLAP_ALLREDUCE=0 run 29522 r2_bench_${N}gpu_nooverlap.log
echo "one GPU, same box"; timeout 600 python bench.py --steps 20 --warmup 5 --no-ddim --no-cpu --no-gpu-ref --no-cfg1 > gpurun_out/r2_bench_1gpu_samebox8.log 2>&1; grep '^{' gpurun_out/r2_bench_1gpu_samebox8.log | tail -1 | cut -c1-220

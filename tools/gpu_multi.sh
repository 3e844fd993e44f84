#!/bin/bash
# N-GPU bench under torchrun (N = $1)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 | tee gpurun_out/topo_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>&1 | grep -E '^\{|Error|error' | tee gpurun_out/bench_n$N.log

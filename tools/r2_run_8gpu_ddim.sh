#!/bin/bash
# BASELINE configs[3]: DDIM-100 do-intervention sweep, 4096 interventions over 8 GPUs (512 per GPU), one gather at the end
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-gpu-ref --no-cfg1 --ddim-steps 100 > gpurun_out/r2_bench_${N}gpu_ddim100.log 2>&1
grep '^{' gpurun_out/r2_bench_${N}gpu_ddim100.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step']); print(d['ddim'])"

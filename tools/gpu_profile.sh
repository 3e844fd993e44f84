#!/bin/bash
mkdir -p gpurun_out
# (1) every launch with its device time for two steady-state steps (skip plan build + graph capture warm-up)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-ddim --no-cpu > gpurun_out/ncu_bench_stdout.log 2>&1
tail -3 gpurun_out/ncu_bench_stdout.log
# (2) full capture of the dominant kernel (3 launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 400 -c 3 -o gpurun_out/prof_igemm \
    python bench.py --steps 2 --warmup 3 --no-ddim --no-cpu > gpurun_out/ncu_full_stdout.log 2>&1
tail -3 gpurun_out/ncu_full_stdout.log
# (3) eager PyTorch reference on the same GPU
timeout 900 python tests/gpu_ref_eager.py 2>&1 | tail -3 | tee gpurun_out/ref_eager.log
ls -la gpurun_out

#!/bin/bash
# ncu --set full captures of the remaining kernel classes (one launch each) for profiles/
mkdir -p gpurun_out
run() { # name regex script args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o gpurun_out/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
}
run wgrad3 wgrad3_kernel python tools/gpu_igemm_bench.py wgrad
run wgrad1 "wgrad_kernel" python tools/gpu_igemm_bench.py wgrad
run attn_fwd attn_fwd_kernel python tools/gpu_attn_bench.py
run attn_dq attn_bwd_dq python tools/gpu_attn_bench.py
run attn_dkv attn_bwd_dkv python tools/gpu_attn_bench.py
run igemm3_256 "igemm3_kernel<256" python tools/gpu_igemm_bench.py fwd
timeout 900 ncu --set full --clock-control none -k regex:"adam_ema|ddim_kernel|q_sample|mse_kernel" -c 6 -o gpurun_out/prof_elementwise -f python bench.py --steps 1 --warmup 3 --no-ddim --no-cpu > gpurun_out/ncu_elementwise.log 2>&1
ls -la gpurun_out | grep prof_

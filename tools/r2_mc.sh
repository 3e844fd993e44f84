#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16; dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
N, H, cin, cout, ks = 64, 8, 512, 512, 3
x = torch.randn(N, H, H, cin, device=dev, generator=g).to(bf16); dy = torch.randn(N, H, H, cout, device=dev, generator=g).to(bf16)
dw = torch.zeros(cout, ks * ks, cin, device=dev)
d = ops.make_wgrad_desc(dy, x, dw, cout, cin, ksize=ks, splits=0)
for _ in range(6): ops.wgrad(d)
torch.cuda.synchronize()
PY
CDAE_WGRAD_CLUSTER=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:wgrad_kernel -s 3 -c 1 -o gpurun_out/r2_wgrad256_8x8 python /tmp/one.py > gpurun_out/ncu_w.log 2>&1; tail -2 gpurun_out/ncu_w.log

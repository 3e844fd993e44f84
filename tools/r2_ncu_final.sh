#!/bin/bash
# ncu --set full of the roofline kernel (3x3 conv 128->128 @64x64, batch 64, with the statistics epilogue) as it is at the end
# of the round, and of the same layer with GroupNorm+SiLU applied on load (sampling form)
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16; dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
N, H, C = 64, 64, 128
segs, K = ops.conv_segments([C], 3)
w = (torch.randn(C, K, device=dev, generator=g) * K ** -0.5).to(bf16)
bias = torch.zeros(C, device=dev)
x = torch.randn(N, H, H, C, device=dev, generator=g).to(bf16)
out = torch.empty(N, H, H, C, device=dev, dtype=bf16)
st = torch.zeros(N, C, 2, device=dev)
xs = torch.stack([x.float().sum(dim=(1, 2)), (x.float() ** 2).sum(dim=(1, 2))], dim=-1).contiguous()
ab = torch.empty(N, C, 2, device=dev)
ops.gn_apply_fwd(x, xs, torch.ones(C, device=dev), torch.zeros(C, device=dev), ab=ab, constants_only=True)
d0 = ops.make_igemm_desc([x], segs, w, out, C, bias=bias, stats=st)
d1 = ops.make_igemm_desc([x], segs, w, out, C, bias=bias, stats=st, gn=(ab, [0]))
for _ in range(4):
    ops.igemm(d0); ops.igemm(d1)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm3t -s 4 -c 2 -o gpurun_out/r2_igemm3t_final python /tmp/one.py > gpurun_out/ncu_f.log 2>&1; tail -2 gpurun_out/ncu_f.log

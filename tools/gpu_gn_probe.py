"""Probe: does GroupNorm throughput depend on the row fraction a unit touches?  Same bytes, different (C, HW, B)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")

def timeit(fns, iters=12):
    for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fns[i % len(fns)]()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

g = torch.Generator(device=dev).manual_seed(0)
for (B, C, HW) in [(256, 32, 4096), (128, 64, 4096), (64, 128, 4096), (32, 256, 4096), (256, 128, 1024), (1024, 128, 256), (512, 256, 256)]:
    gamma, beta = torch.randn(C, device=dev, generator=g), torch.randn(C, device=dev, generator=g)
    ff, fc = [], []
    for _ in range(3):
        x0 = torch.randn(B, HW, 1, C, device=dev, generator=g).to(bf16)
        y = torch.empty_like(x0)
        mean, rstd = torch.empty(B, 32, device=dev), torch.empty(B, 32, device=dev)
        ff.append(lambda x0=x0, y=y, mean=mean, rstd=rstd: ops.gn_fwd(x0, gamma, beta, silu=True, out=y, mean=mean, rstd=rstd))
        fc.append(lambda x0=x0, y=y: y.copy_(x0))
    tf, tc = timeit(ff), timeit(fc)
    n = B * HW * C
    print(f"B {B:4d} C {C:4d} HW {HW:5d}: gn_fwd {tf*1e3:7.1f} us {4*n/tf/1e6:6.0f} GB/s | torch copy {tc*1e3:7.1f} us {4*n/tc/1e6:6.0f} GB/s", flush=True)

"""Attention micro-benchmark on the cfg2 shapes (batch 64): per-launch time and achieved tensor throughput
(forward 4*T^2*ch FLOP per (batch, head); backward counted as 2.5x forward = the algorithmic five GEMMs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
B = 64

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

g = torch.Generator(device=dev).manual_seed(0)
for (T, heads, ch, cnt) in [(256, 4, 96, 5), (64, 4, 128, 6), (1024, 4, 64, 0), (16, 4, 64, 0)]:
    C = heads * ch
    qkv = torch.randn(B, T, 3 * C, device=dev, generator=g).to(bf16)
    dout = torch.randn(B, T, C, device=dev, generator=g).to(bf16)
    out, lse = ops.attn_fwd(qkv, heads)
    dqkv = torch.empty_like(qkv); dsum = torch.empty(B, heads, T, device=dev)
    tf = timeit(lambda: ops.attn_fwd(qkv, heads, out=out, lse=lse))
    tb = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, heads, dqkv=dqkv, dsum=dsum))
    fl = 4.0 * T * T * ch * B * heads
    print(f"T {T:5d} heads {heads} ch {ch:3d} x{cnt}: fwd {tf*1e3:7.1f} us {fl/tf/1e9:6.1f} TF/s | bwd {tb*1e3:7.1f} us {2.5*fl/tb/1e9:6.1f} TF/s", flush=True)

"""Sweep of the split-K factor of the weight-gradient kernels on the small / 1x1 cfg2 layers (batch 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
SHAPES = [(64, 64, 128, 128, 1), (64, 32, 256, 256, 1), (64, 16, 384, 384, 1), (64, 16, 384, 1152, 1), (64, 8, 512, 512, 1),
          (64, 8, 512, 1536, 1), (64, 8, 1024, 512, 1), (64, 8, 512, 512, 3), (64, 16, 384, 384, 3), (64, 32, 256, 256, 3),
          (64, 64, 128, 128, 3)]


def timeit(fns, iters=12):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator(device=dev).manual_seed(0)
for (N, H, cin, cout, ks) in SHAPES:
    res = []
    sets = [(torch.randn(N, H, H, cin, device=dev, generator=g).to(bf16), torch.randn(N, H, H, cout, device=dev, generator=g).to(bf16))
            for _ in range(3)]
    dw = torch.zeros(cout, ks * ks, cin, device=dev)
    for sp in (0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48):
        fns = [(lambda d=ops.make_wgrad_desc(dy, x, dw, cout, cin, ksize=ks, splits=sp): ops.wgrad(d)) for x, dy in sets]
        try:
            res.append((sp, timeit(fns) * 1e3))
        except Exception as ex:
            res.append((sp, float("nan")))
    fl = 2.0 * N * H * H * cout * ks * ks * cin
    best = min(res[1:], key=lambda t: t[1])
    print(f"wgrad N{N} {H}x{H} cin{cin} cout{cout} k{ks}: " + " ".join(f"s{sp}:{t:.1f}" for sp, t in res) +
          f" | auto {res[0][1]:.1f} us, best s{best[0]} {best[1]:.1f} us ({fl / best[1] / 1e6:.0f} TF/s)", flush=True)

"""Weight-gradient kernels on the cfg2 shapes that take the one-tap kernel (cin % 256 == 0, 3x3; and the 1x1 layers), isolated:
   python tools/r2_wgrad_mc.py    (run once per CDAE_WGRAD_CLUSTER value: the cluster size is read once per process)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
# (N, H, cin, cout, ksize)
SHAPES = [(64, 8, 512, 512, 3), (64, 8, 1024, 512, 3), (64, 16, 768, 384, 3), (64, 16, 256, 384, 3), (64, 32, 256, 256, 3),
          (64, 32, 512, 256, 3), (64, 64, 256, 128, 3), (64, 8, 512, 512, 1), (64, 8, 512, 1536, 1), (64, 16, 768, 384, 1),
          (64, 32, 256, 256, 1)]


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
tot = 0.0
for (N, H, cin, cout, ks) in SHAPES:
    sets = [(torch.randn(N, H, H, cin, device=dev, generator=g).to(bf16), torch.randn(N, H, H, cout, device=dev, generator=g).to(bf16))
            for _ in range(3)]
    dw = torch.zeros(cout, ks * ks, cin, device=dev)
    fns = [(lambda d=ops.make_wgrad_desc(dy, x, dw, cout, cin, ksize=ks, splits=0): ops.wgrad(d)) for x, dy in sets]
    t = timeit(fns) * 1e3
    fl = 2.0 * N * H * H * cout * ks * ks * cin
    tot += t
    print(f"wgrad N{N} {H}x{H} cin{cin} cout{cout} k{ks}: {t:7.1f} us {fl / t / 1e6:6.0f} TF/s", flush=True)
print(f"sum {tot:.1f} us (cluster setting {os.environ.get('CDAE_WGRAD_CLUSTER', 'default 4')})")

"""N-tile sweep of the conv kernel on the small-spatial cfg2 layers (batch 64): few, long-K tiles are bound by L2 -> SM traffic,
which a wider N tile cuts even when it leaves SMs idle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
SHAPES = [(64, 8, [512], 512, 3), (64, 8, [512, 512], 512, 3), (64, 8, [512, 384], 512, 3), (64, 16, [384], 384, 3),
          (64, 16, [384, 384], 384, 3), (64, 16, [256], 384, 3), (64, 8, [384], 512, 3), (64, 16, [384], 1152, 1), (64, 8, [512], 1536, 1),
          (64, 32, [256], 256, 3)]


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
for (N, H, chans, cout, ks) in SHAPES:
    cin = sum(chans)
    segs, K = ops.conv_segments(chans, ks)
    w = (torch.randn(cout, K, device=dev, generator=g) * 0.02).to(bf16)
    bias = torch.zeros(cout, device=dev)
    sets = []
    for _ in range(3):
        xs = [torch.randn(N, H, H, c, device=dev, generator=g).to(bf16) for c in chans]
        sets.append((xs, torch.empty(N, H, H, cout, device=dev, dtype=bf16)))
    fl = 2.0 * N * H * H * cout * K
    line = f"conv N{N} {H}x{H} cin{chans} cout{cout} k{ks}:"
    for bn in (0, 64, 128, 192, 256):
        if bn and cout % bn:
            continue
        try:
            fns = [(lambda d=ops.make_igemm_desc(xs, segs, w, out, cout, bias=bias, bn=bn): ops.igemm(d)) for xs, out in sets]
            t = timeit(fns) * 1e3
            line += f"  bn{bn}: {t:6.1f} us ({fl / t / 1e6:5.0f} TF/s)"
        except Exception as ex:
            line += f"  bn{bn}: n/a"
    print(line, flush=True)

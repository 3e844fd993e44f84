"""Micro-benchmark of the tensor-core conv kernels on the cfg2 shapes (batch 64): igemm forward / wgrad, CUDA events,
three rotating buffer sets (> 126 MB L2 for the big layers).  Prints TFLOP/s and the fraction of the measured bf16 peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops

bf16 = torch.bfloat16
dev = torch.device("cuda:0")
PEAK = 1677.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass

# (N, H, chans, cout, ksize, stride, resid)
FWD = [
    (64, 64, [128], 128, 3, 1, False), (64, 64, [128], 128, 3, 1, True), (64, 64, [256], 256, 3, 1, False),
    (64, 64, [128, 128], 128, 3, 1, False), (64, 32, [256], 256, 3, 1, True), (64, 32, [384, 256], 256, 3, 1, False),
    (64, 16, [384], 384, 3, 1, True), (64, 8, [512], 512, 3, 1, True), (64, 8, [512, 512], 512, 3, 1, False),
    (64, 64, [128], 128, 1, 1, True), (64, 16, [384], 1152, 1, 1, False), (64, 16, [384], 384, 1, 1, True),
    (64, 8, [512], 512, 1, 1, True), (64, 64, [128], 128, 3, 2, False),
    (64, 64, [256], 256, 3, 1, False, 128), (64, 32, [256], 256, 3, 1, True, 128), (64, 16, [384], 384, 3, 1, True, 128),
    (64, 32, [384, 256], 256, 3, 1, False, 128), (64, 64, [128], 3, 3, 1, False),
]
WG = [(64, 64, 128, 128, 3), (64, 32, 256, 256, 3), (64, 16, 384, 384, 3), (64, 8, 512, 512, 3), (64, 64, 256, 128, 3),
      (64, 64, 128, 128, 1), (64, 16, 384, 384, 1)]


def timeit(fns, iters=12):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    with_stats = "stats" in sys.argv[2:]     # also accumulate the GroupNorm channel sums in the epilogue
    g = torch.Generator(device=dev).manual_seed(0)
    if only in ("", "fwd"):
        for case in (FWD[:1] if "first" in sys.argv[2:] else FWD):
            (N, H, chans, cout, ks, st, resid), bn = case[:7], (case[7] if len(case) > 7 else 0)
            cin = sum(chans)
            OH = H // st
            w = (torch.randn(max(cout, 16), ks * ks * cin, device=dev, generator=g) * 0.02).to(bf16)
            bias = torch.randn(cout, device=dev, generator=g)
            fns = []
            for _ in range(3):
                xs = [torch.randn(N, H, H, c, device=dev, generator=g).to(bf16) for c in chans]
                mode1 = cout < 8
                out = torch.empty(N, cout, OH, OH, device=dev) if mode1 else torch.empty(N, OH, OH, cout, device=dev, dtype=bf16)
                r = torch.randn(N, OH, OH, cout, device=dev, generator=g).to(bf16) if resid else None
                segs, K = ops.conv_segments(chans, ks)
                stt = torch.zeros(N, cout, 2, device=dev) if (with_stats and not mode1 and cout % 64 == 0) else None
                d = ops.make_igemm_desc(xs, segs, w, out, cout, in_stride=st, bias=bias, resid=r, bn=bn,
                                        out_mode=1 if mode1 else 0, stats=stt)
                fns.append(lambda d=d: ops.igemm(d))
            ms = timeit(fns)
            fl = 2.0 * N * OH * OH * cout * ks * ks * cin
            byts = 2.0 * N * (H * H * cin + OH * OH * cout * (2 if resid else 1))
            print(f"igemm{'+stats' if with_stats else ''} N{N} {H}x{H} cin{chans} cout{cout} k{ks} s{st} resid{int(resid)} bn{bn}: {ms*1e3:8.1f} us "
                  f"{fl/ms/1e9:7.1f} TF/s ({fl/ms/1e9/PEAK:5.1%} of measured peak)  {byts/ms/1e6:6.0f} GB/s algorithmic", flush=True)
    if only in ("", "wgrad"):
        for (N, H, cin, cout, ks) in WG:
            fns = []
            dw = torch.zeros(cout, ks * ks, cin, device=dev)
            for _ in range(3):
                x = torch.randn(N, H, H, cin, device=dev, generator=g).to(bf16)
                dy = torch.randn(N, H, H, cout, device=dev, generator=g).to(bf16)
                d = ops.make_wgrad_desc(dy, x, dw, cout, cin, ksize=ks)
                fns.append(lambda d=d: ops.wgrad(d))
            ms = timeit(fns)
            fl = 2.0 * N * H * H * cout * ks * ks * cin
            print(f"wgrad N{N} {H}x{H} cin{cin} cout{cout} k{ks}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s "
                  f"({fl/ms/1e9/PEAK:5.1%} of measured peak)", flush=True)


if __name__ == "__main__":
    main()

"""3x3 conv with GroupNorm+SiLU on load (inference) against the two-pass form, isolated, at sampling batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
SHAPES = [(512, 64, [128], 128), (512, 64, [128, 128], 128), (512, 32, [256], 256), (512, 32, [256, 256], 256), (512, 64, [256, 128], 128)]


def timeit(fns, iters=8):
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
for (N, H, chans, cout) in SHAPES:
    Ct = sum(chans)
    segs, K = ops.conv_segments(chans, 3)
    w = (torch.randn(cout, K, device=dev, generator=g) * K ** -0.5).to(bf16)
    bias = torch.zeros(cout, device=dev)
    gamma, beta = torch.ones(Ct, device=dev), torch.zeros(Ct, device=dev)
    xs = [torch.randn(N, H, H, c, device=dev, generator=g).to(bf16) for c in chans]
    sts = [torch.stack([x.float().sum(dim=(1, 2)), (x.float() ** 2).sum(dim=(1, 2))], dim=-1).contiguous() for x in xs]
    a = torch.empty(N, H, H, Ct, device=dev, dtype=bf16)
    out = torch.empty(N, H, H, cout, device=dev, dtype=bf16)
    ab = torch.empty(N, Ct, 2, device=dev)
    x1, s1 = (xs[1], sts[1]) if len(xs) > 1 else (None, None)
    f_apply = lambda: ops.gn_apply_fwd(xs[0], sts[0], gamma, beta, x1=x1, stats1=s1, out=a)
    a_srcs = [a[..., :chans[0]].contiguous()] + ([a[..., chans[0]:].contiguous()] if x1 is not None else [])
    d_plain = ops.make_igemm_desc(a_srcs, segs, w, out, cout, bias=bias)
    f_conv = lambda: ops.igemm(d_plain)
    f_const = lambda: ops.gn_apply_fwd(xs[0], sts[0], gamma, beta, x1=x1, stats1=s1, ab=ab, constants_only=True)
    offs = [0] + ([chans[0]] if x1 is not None else [])
    d_fused = ops.make_igemm_desc(xs, segs, w, out, cout, bias=bias, gn=(ab, offs))
    f_fused = lambda: ops.igemm(d_fused)
    fl = 2.0 * N * H * H * cout * K
    ta, tc, tk, tf = timeit([f_apply]), timeit([f_conv]), timeit([f_const]), timeit([f_fused])
    print(f"N{N} {H}x{H} cin{chans} cout{cout}: apply {ta*1e3:7.1f} us + conv {tc*1e3:7.1f} us ({fl/tc/1e9:5.0f} TF/s) = {1e3*(ta+tc):7.1f} | "
          f"constants {tk*1e3:5.1f} us + fused conv {tf*1e3:7.1f} us ({fl/tf/1e9:5.0f} TF/s) = {1e3*(tk+tf):7.1f}  ({(ta+tc)/(tk+tf):.3f}x)", flush=True)

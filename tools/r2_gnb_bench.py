"""Round-2 micro-benchmark of the fused GroupNorm backward: the data-gradient conv with / without the gnb epilogue and
the norm's backward as resident kernel vs streaming apply pass, per cfg2 layer shape (batch 64), CUDA events, rotating
buffers larger than L2."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops

bf16 = torch.bfloat16
dev = torch.device("cuda:0")
P = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"bf16_tflops": 1640.0, "hbm_gbs": 6543.0}
# (N, H, C0, C1, conv_cout, ksize): norm over C0+C1 channels feeding a conv (C0+C1 -> conv_cout)
SHAPES = [(64, 64, 128, 0, 128, 3), (64, 64, 128, 128, 128, 3), (64, 32, 256, 0, 256, 3), (64, 16, 384, 0, 384, 3),
          (64, 8, 512, 0, 512, 3), (64, 16, 384, 0, 1152, 1)]
if "first" in sys.argv[1:]:
    SHAPES = SHAPES[:1]


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
for (N, H, C0, C1, cc, ks) in SHAPES:
    Ct = C0 + C1
    gamma, beta = torch.randn(Ct, device=dev, generator=g), torch.randn(Ct, device=dev, generator=g)
    wt = (torch.randn(Ct, ks * ks * cc, device=dev, generator=g) * 0.02).to(bf16)
    segs, _ = ops.conv_segments([cc], ks, transposed=True)
    plain, fused, res, app = [], [], [], []
    for _ in range(3):
        x0 = torch.randn(N, H, H, C0, device=dev, generator=g).to(bf16)
        x1 = torch.randn(N, H, H, C1, device=dev, generator=g).to(bf16) if C1 else None
        dz = torch.randn(N, H, H, cc, device=dev, generator=g).to(bf16)
        du = torch.empty(N, H, H, Ct, device=dev, dtype=bf16)
        ab = torch.randn(N, Ct, 2, device=dev, generator=g) * 0.3
        ws = torch.zeros(N, Ct, 2, device=dev)
        mean, rstd = torch.randn(N, 32, device=dev, generator=g) * 0.1, torch.rand(N, 32, device=dev, generator=g) + 0.5
        dx0 = torch.empty_like(x0); dx1 = torch.empty_like(x1) if C1 else None
        dg, db = torch.zeros(Ct, device=dev), torch.zeros(Ct, device=dev)
        d0 = ops.make_igemm_desc([dz], segs, wt, du, Ct)
        d1 = ops.make_igemm_desc([dz], segs, wt, du, Ct, gnb=dict(x0=x0, x1=x1, ab=ab, ws=ws, silu=True))
        plain.append(lambda d=d0: ops.igemm(d))
        fused.append(lambda d=d1: ops.igemm(d))
        res.append(lambda du=du, x0=x0, x1=x1, mean=mean, rstd=rstd, dx0=dx0, dx1=dx1: ops.gn_bwd(
            du, x0, gamma, beta, mean, rstd, x1=x1, silu=True, dx0=dx0, dx1=dx1, dgamma=dg, dbeta=db))
        app.append(lambda du=du, x0=x0, x1=x1, mean=mean, rstd=rstd, dx0=dx0, dx1=dx1, ws=ws: ops.gn_bwd_apply(
            du, x0, gamma, beta, mean, rstd, ws, x1=x1, dx0=dx0, dx1=dx1, dgamma=dg, dbeta=db))
    t0, t1, t2, t3 = timeit(plain), timeit(fused), timeit(res), timeit(app)
    fl = 2.0 * N * H * H * Ct * ks * ks * cc
    by = 6.0 * N * H * H * Ct
    print(f"N{N} {H}x{H} C{C0}+{C1} conv->{cc} k{ks}: dgrad {t0*1e3:7.1f} us ({fl/t0/1e9:6.0f} TF/s) | dgrad+gnb {t1*1e3:7.1f} us "
          f"({fl/t1/1e9:6.0f} TF/s, +{(t1-t0)*1e3:5.1f}) | gn_bwd resident {t2*1e3:7.1f} us ({by/t2/1e6/P['hbm_gbs']:.2f} of HBM) | "
          f"apply {t3*1e3:7.1f} us ({by/t3/1e6/P['hbm_gbs']:.2f}) | net {(t1+t3-t0-t2)*1e3:+6.1f} us", flush=True)

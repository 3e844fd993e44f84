"""GroupNorm micro-benchmark on the cfg2 shapes (batch 64): per-shape time and achieved algorithmic HBM GB/s
(forward 4 B/element, backward 6 B/element + 2 B per accumulated/added tensor), weighted per-step totals."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from causaldiffae_b200 import ops
bf16 = torch.bfloat16
dev = torch.device("cuda:0")
PEAK = 6543.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
# (C0, C1, HW, count per forward)  -- SURVEY appendix A.1
SHAPES = [(128, 0, 4096, 8), (128, 0, 1024, 1), (256, 0, 1024, 6), (256, 0, 256, 1), (384, 0, 256, 11), (384, 0, 64, 1),
          (512, 0, 64, 16), (512, 512, 64, 2), (512, 384, 64, 1), (512, 384, 256, 1), (384, 384, 256, 1), (384, 256, 256, 1),
          (384, 256, 1024, 1), (256, 256, 1024, 1), (256, 128, 1024, 1), (256, 128, 4096, 1), (128, 128, 4096, 2)]
B = 64
if 'first' in sys.argv[1:]:
    SHAPES = SHAPES[:1]     # ncu captures: the 128-channel 64x64 layer only

def timeit(fns, iters=12):
    for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fns[i % len(fns)]()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

tot_f = tot_b = tot_a = tot_s = ideal_f = ideal_b = 0.0
g = torch.Generator(device=dev).manual_seed(0)
for (C0, C1, HW, cnt) in SHAPES:
    C = C0 + C1
    gamma, beta = torch.randn(C, device=dev, generator=g), torch.randn(C, device=dev, generator=g)
    film = torch.randn(B, 2 * C, device=dev, generator=g) * 0.1
    ff, fb, fa, fs = [], [], [], []
    for _ in range(3):
        x0 = torch.randn(B, HW, 1, C0, device=dev, generator=g).to(bf16)
        x1 = torch.randn(B, HW, 1, C1, device=dev, generator=g).to(bf16) if C1 else None
        y = torch.empty(B, HW, 1, C, device=dev, dtype=bf16)
        mean, rstd = torch.empty(B, 32, device=dev), torch.empty(B, 32, device=dev)
        dy = torch.randn(B, HW, 1, C, device=dev, generator=g).to(bf16)
        dx0 = torch.empty_like(x0); dx1 = torch.empty_like(x1) if C1 else None
        dg, db, dfilm = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros_like(film)
        st0 = torch.stack([x0.float().sum(dim=(1, 2)), (x0.float() ** 2).sum(dim=(1, 2))], dim=-1).contiguous()
        st1 = torch.stack([x1.float().sum(dim=(1, 2)), (x1.float() ** 2).sum(dim=(1, 2))], dim=-1).contiguous() if C1 else None
        fa.append(lambda x0=x0, x1=x1, y=y, mean=mean, rstd=rstd, st0=st0, st1=st1: ops.gn_apply_fwd(x0, st0, gamma, beta, x1=x1, stats1=st1, film=film, silu=True, out=y, mean=mean, rstd=rstd))
        ff.append(lambda x0=x0, x1=x1, y=y, mean=mean, rstd=rstd: ops.gn_fwd(x0, gamma, beta, x1=x1, film=film, silu=True, out=y, mean=mean, rstd=rstd))
        ws = torch.randn(B, C, 2, device=dev, generator=g)      # {sum du, sum du*x} as the data-gradient epilogue leaves them
        fs.append(lambda x0=x0, x1=x1, dy=dy, mean=mean, rstd=rstd, dx0=dx0, dx1=dx1, ws=ws: ops.gn_bwd_apply(dy, x0, gamma, beta, mean, rstd, ws, x1=x1, film=film, dx0=dx0, dx1=dx1, dgamma=dg, dbeta=db, dfilm=dfilm))
        fb.append(lambda x0=x0, x1=x1, dy=dy, mean=mean, rstd=rstd, dx0=dx0, dx1=dx1: ops.gn_bwd(dy, x0, gamma, beta, mean, rstd, x1=x1, film=film, silu=True, dx0=dx0, dx1=dx1, dgamma=dg, dbeta=db, dfilm=dfilm))
    tf = timeit(ff); tb = timeit(fb); ta = timeit(fa); ts = timeit(fs)
    n = B * HW * C
    print(f"C {C0:4d}+{C1:4d} HW {HW:5d} x{cnt:2d}: apply {ta*1e3:7.1f} us {4*n/ta/1e6:6.0f} GB/s ({4*n/ta/1e6/PEAK:5.1%}) | fwd {tf*1e3:7.1f} us {4*n/tf/1e6:6.0f} GB/s ({4*n/tf/1e6/PEAK:5.1%}) | "
          f"bwd {tb*1e3:7.1f} us {6*n/tb/1e6:6.0f} GB/s ({6*n/tb/1e6/PEAK:5.1%}) | bwd-apply {ts*1e3:7.1f} us {6*n/ts/1e6:6.0f} GB/s ({6*n/ts/1e6/PEAK:5.1%})", flush=True)
    tot_a += cnt * ta; tot_s += cnt * ts; tot_f += cnt * tf; tot_b += cnt * tb; ideal_f += cnt * 4 * n / PEAK / 1e6; ideal_b += cnt * 6 * n / PEAK / 1e6
print(f"per step: streaming fwd (stats from the conv epilogue) {tot_a:.3f} ms ({ideal_f/tot_a:5.1%} of HBM-ideal)")
print(f"per step: streaming bwd apply pass (statistics from the data-gradient epilogue, opt-in) {tot_s:.3f} ms ({ideal_b/tot_s:5.1%} of HBM-ideal)")
print(f"per step: fwd {tot_f:.3f} ms (HBM-ideal {ideal_f:.3f}, {ideal_f/tot_f:5.1%}) | bwd {tot_b:.3f} ms (ideal {ideal_b:.3f}, {ideal_b/tot_b:5.1%})")

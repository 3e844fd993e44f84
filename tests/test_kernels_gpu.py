"""Kernel-level parity: every libcdae entry point (through the C ABI) against plain fp32 torch / the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


def dev():
    return torch.device("cuda:0")


def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def pack_ohwi(w):
    """OIHW fp32 -> bf16 [Cout, taps*Cin] (tap-major, channel-minor)"""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(bf16)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(bf16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------ diffusion elementwise
def test_q_sample_bit_exact_vs_oracle():
    from causaldiffae_b200 import ops
    from oracle import diffusion as od
    d = od.Diffusion(steps=1000)
    g = torch.Generator().manual_seed(0)
    B = 37
    x0, nz = torch.rand(B, 3, 64, 64, generator=g), torch.randn(B, 3, 64, 64, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    ref = d.q_sample(x0, t, nz)
    ta = torch.from_numpy(d.tables["sqrt_alphas_cumprod"]).float().to(dev())
    tb = torch.from_numpy(d.tables["sqrt_one_minus_alphas_cumprod"]).float().to(dev())
    out = ops.q_sample(x0.to(dev()), nz.to(dev()), t.to(dev()), ta, tb)
    assert torch.equal(out.cpu(), ref)
    # empty batch is a no-op
    e = ops.q_sample(x0[:0].to(dev()), nz[:0].to(dev()), t[:0].to(dev()), ta, tb)
    assert e.shape[0] == 0


def test_mse_loss_and_grad():
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(1)
    B = 9
    pred, tgt = torch.randn(B, 3, 32, 32, generator=g), torch.randn(B, 3, 32, 32, generator=g)
    w = torch.rand(B, generator=g)
    p = pred.clone().requires_grad_(True)
    mse_ref = ((tgt - p) ** 2).mean(dim=(1, 2, 3))
    (mse_ref * w).sum().backward()
    mse, dp = ops.mse_loss(pred.to(dev()), tgt.to(dev()), w.to(dev()), want_grad=True)
    np.testing.assert_allclose(mse.cpu().numpy(), mse_ref.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(dp.cpu().numpy(), p.grad.numpy(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("eta,w", [(0.0, None), (0.0, 2.0), (0.7, 1.5)])
def test_ddim_step_vs_oracle(eta, w):
    from causaldiffae_b200 import ops
    from causaldiffae_b200.respace import ddim_coef_table
    from oracle import diffusion as od
    d = od.Diffusion(steps=1000, timestep_respacing="ddim50")
    g = torch.Generator().manual_seed(2)
    B = 5
    x, ec, eu, nz = (torch.randn(B, 3, 64, 64, generator=g) for _ in range(4))
    tab = torch.from_numpy(ddim_coef_table(d.tables, eta=eta, clip_denoised=True)).to(dev())
    for ti in (49, 17, 0):
        t = torch.full((B,), ti, dtype=torch.long)
        ref, x0ref = d.ddim_step(x, t, ec, eu if w is not None else None, w, eta, nz)
        out, x0 = ops.ddim_step(x.to(dev()), ec.to(dev()), tab, torch.tensor([ti], dtype=torch.int32, device=dev()),
                                eps_u=eu.to(dev()) if w is not None else None, w=w, noise=nz.to(dev()), want_xstart=True)
        np.testing.assert_allclose(x0.cpu().numpy(), x0ref.numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-6, atol=2e-6)
    # per-sample timestep indices
    tt = torch.tensor([0, 3, 49, 20, 7])
    ref, _ = d.ddim_step(x, tt, ec, None, None, eta, nz)
    out, _ = ops.ddim_step(x.to(dev()), ec.to(dev()), tab, tt.to(torch.int32).to(dev()), noise=nz.to(dev()))
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-6, atol=2e-6)


def test_adam_ema_vs_torch_adamw():
    """fused AdamW + EMA + sum(g^2) against torch.optim.AdamW (ref train_util.py:292-303): device step counter, device-side
    bias corrections, the non-finite guard of optimize_fp16 (ref :277-280) and bf16 gradients (the all-reduced copy)"""
    from causaldiffae_b200 import ops
    from causaldiffae_b200.train_util import adam_hyper
    g = torch.Generator().manual_seed(3)
    n = 4096 * 3 + 4
    p0 = torch.randn(n, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=0.01)
    ema_ref = p0.clone()
    p, m, v, ema = p0.clone().to(dev()), torch.zeros(n, device=dev()), torch.zeros(n, device=dev()), p0.clone().to(dev())
    gsq, guard = torch.zeros(1, device=dev()), torch.zeros(1, device=dev())
    step = torch.zeros(1, device=dev(), dtype=torch.int64)
    hyper = torch.tensor(adam_hyper(1e-3, weight_decay=0.01, ema_rate=0.99) + [0.0], device=dev())
    for it in range(1, 6):
        grad = torch.randn(n, generator=g)
        pr.grad = grad.clone()
        opt.step()
        ema_ref.mul_(0.99).add_(pr.detach(), alpha=0.01)
        gsq.zero_(); guard.zero_()
        gd = grad.to(dev())
        ops.sumsq(gd, guard)
        ops.adam_ema(p, gd, m, v, ema, hyper, step, gsq, guard)
        np.testing.assert_allclose(float(gsq), float((grad ** 2).sum()), rtol=1e-5)
        np.testing.assert_allclose(float(guard), float((grad ** 2).sum()), rtol=1e-5)
        assert int(step) == it
    np.testing.assert_allclose(p.cpu().numpy(), pr.detach().numpy(), rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(ema.cpu().numpy(), ema_ref.numpy(), rtol=2e-5, atol=2e-7)
    # a NaN / Inf gradient: the guarded launch changes nothing and the step counter stays
    snap = [t.clone() for t in (p, m, v, ema)]
    for bad in (float("nan"), float("inf")):
        gd = torch.randn(n, generator=g).to(dev()); gd[n // 2] = bad
        guard.zero_(); gsq.zero_()
        ops.sumsq(gd, guard)
        ops.adam_ema(p, gd, m, v, ema, hyper, step, gsq, guard)
        assert int(step) == 5 and not bool(torch.isfinite(guard).all())
        for a, b in zip(snap, (p, m, v, ema)):
            assert torch.equal(a, b)
    # bf16 gradients (as they come off the all-reduce): same step as fp32 Adam on the rounded values
    gd = torch.randn(n, generator=g).to(dev())
    gb = ops.cast_bf16(gd)
    assert torch.equal(gb, gd.to(bf16))
    p2, m2, v2, e2, s2 = p.clone(), m.clone(), v.clone(), ema.clone(), step.clone()
    ops.adam_ema(p, gb, m, v, ema, hyper, step, None, None)
    ops.adam_ema(p2, gb.float(), m2, v2, e2, hyper, s2, None, None)
    assert torch.equal(p, p2) and torch.equal(m, m2) and torch.equal(v, v2) and torch.equal(ema, e2) and int(step) == 6


# ------------------------------------------------------------------ layout kernels
def test_layout_kernels():
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 3, 16, 16, generator=g).to(dev())
    y = ops.nchw_to_nhwc_pad(x, 64)
    assert y.shape == (3, 16, 16, 64)
    assert torch.equal(y[..., :3].float(), x.permute(0, 2, 3, 1).to(bf16).float()) and float(y[..., 3:].abs().max()) == 0
    assert torch.equal(ops.nhwc_to_nchw(y, 3), x.to(bf16).float())
    a = torch.randn(2, 8, 8, 64, generator=g).to(dev()).to(bf16)
    up = ops.upsample2x(a)
    assert torch.equal(nchw(up), F.interpolate(nchw(a), scale_factor=2, mode="nearest"))
    sp = ops.sumpool2x(up)
    assert relerr(sp, 4 * a.float()) < 1e-2
    zi = ops.zero_insert2x(a)
    assert torch.equal(zi[:, ::2, ::2], a) and float(zi[:, 1::2].abs().max()) == 0 and float(zi[:, :, 1::2].abs().max()) == 0
    cs = torch.zeros(64, device=dev())
    ops.colsum_(a, cs)
    np.testing.assert_allclose(cs.cpu().numpy(), a.float().sum(dim=(0, 1, 2)).cpu().numpy(), rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------ GroupNorm
@pytest.mark.parametrize("C0,C1,HW,film,silu", [(128, 0, 64 * 64, True, True), (64, 0, 16 * 16, False, True),
                                                (512, 384, 8 * 8, False, True), (256, 128, 32 * 32, True, True),
                                                (384, 0, 16 * 16, False, False), (64, 64, 4 * 4, True, True)])
def test_groupnorm_fwd_bwd(C0, C1, HW, film, silu):
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(C0 + C1 + HW)
    B, Ct = 3, C0 + C1
    S = int(HW ** 0.5)
    x = (torch.randn(B, Ct, S, S, generator=g) * 1.5 + 0.3).to(dev()).to(bf16).float()
    gamma = (1 + 0.2 * torch.randn(Ct, generator=g)).to(dev())
    beta = (0.2 * torch.randn(Ct, generator=g)).to(dev())
    film_t = (0.3 * torch.randn(B, 2 * Ct + 32, generator=g)).to(dev()) if film else None
    foff = 16
    dy = torch.randn(B, Ct, S, S, generator=g).to(dev()).to(bf16).float()
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    fr = film_t.clone().requires_grad_(True) if film else None
    u = F.group_norm(xr, 32, gr, br, eps=1e-5)
    if film:
        u = u * (1 + fr[:, foff:foff + Ct, None, None]) + fr[:, foff + Ct:foff + 2 * Ct, None, None]
    yref = u * torch.sigmoid(u) if silu else u
    yref.backward(dy)
    x_nhwc = nhwc(x)
    x0 = x_nhwc[..., :C0].contiguous()
    x1 = x_nhwc[..., C0:].contiguous() if C1 else None
    y, mean, rstd = ops.gn_fwd(x0, gamma, beta, x1=x1, film=film_t, film_off=foff, silu=silu)
    assert relerr(nchw(y), yref) < 6e-3
    dgamma, dbeta = torch.zeros(Ct, device=dev()), torch.zeros(Ct, device=dev())
    dfilm = torch.zeros_like(film_t) if film else None
    dx0, dx1 = ops.gn_bwd(nhwc(dy), x0, gamma, beta, mean, rstd, x1=x1, film=film_t, film_off=foff, silu=silu,
                          dgamma=dgamma, dbeta=dbeta, dfilm=dfilm)
    dx = torch.cat([dx0, dx1], dim=-1) if C1 else dx0
    assert relerr(nchw(dx), xr.grad) < 8e-3
    assert relerr(dgamma, gr.grad) < 3e-3 and relerr(dbeta, br.grad) < 3e-3
    if film:
        assert relerr(dfilm, fr.grad) < 3e-3
    # accumulate mode adds to what is already in dx
    base = torch.randn_like(dx0.float()).to(bf16)
    acc0 = base.clone()
    acc1 = torch.zeros_like(dx1) if C1 else None
    dadd = torch.randn(B, S, S, Ct, generator=g).to(dev()).to(bf16)
    ops.gn_bwd(nhwc(dy), x0, gamma, beta, mean, rstd, x1=x1, film=film_t, film_off=foff, silu=silu, dx0=acc0, dx1=acc1,
               accumulate_dx=True, dadd=dadd)
    assert relerr(acc0.float(), base.float() + dx0.float() + dadd[..., :C0].float()) < 1e-2
    if C1:
        assert relerr(acc1.float(), dx1.float() + dadd[..., C0:].float()) < 1e-2


# ------------------------------------------------------------------ implicit GEMM (tcgen05)
def _conv_case(N, H, chans, cout, ksize=3, stride=1, bias=True, resid=False, seed=0, bn=0):
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(seed)
    cin = sum(chans)
    x = torch.randn(N, cin, H, H, generator=g).to(dev()).to(bf16).float()
    w = (torch.randn(cout, cin, ksize, ksize, generator=g) * (cin * ksize * ksize) ** -0.5).to(dev()).to(bf16).float()
    b = torch.randn(cout, generator=g).to(dev()) if bias else None
    ref = F.conv2d(x, w, b, stride=stride, padding=ksize // 2)
    OH = ref.shape[2]
    r = torch.randn(N, cout, OH, OH, generator=g).to(dev()).to(bf16).float() if resid else None
    if resid:
        ref = ref + r
    xs, off = [], 0
    xn = nhwc(x)
    for c in chans:
        xs.append(xn[..., off:off + c].contiguous()); off += c
    segs, K = ops.conv_segments(chans, ksize)
    out = torch.full((N, OH, OH, cout), float("nan"), device=dev(), dtype=bf16)
    d = ops.make_igemm_desc(xs, segs, pack_ohwi(w), out, cout, in_stride=stride, bias=b,
                            resid=nhwc(r) if resid else None, bn=bn)
    ops.igemm(d)
    torch.cuda.synchronize()
    return relerr(nchw(out), ref)


@pytest.mark.parametrize("N,H,chans,cout,ksize,stride,resid", [
    (2, 64, [64], 128, 3, 1, False),          # BW=64,BH=2
    (2, 32, [128], 128, 3, 1, True),          # residual epilogue
    (3, 16, [128, 64], 128, 3, 1, False),     # concat (split-K over two tensors)
    (5, 8, [256], 256, 3, 1, False),          # 8x8: two images per 128-pixel tile, ragged batch (5)
    (9, 4, [128], 128, 3, 1, False),          # 4x4: eight images per tile, ragged
    (2, 32, [128], 128, 3, 2, False),         # stride-2 (Downsample) via TMA element strides
    (2, 16, [192], 64, 1, 1, True),           # 1x1 skip conv, cout 64
    (1, 64, [64], 64, 3, 1, False),
    (2, 48, [64, 128], 192, 3, 1, True),      # halo kernel: non-power-of-two image, concat, residual, N tile 192
    (3, 24, [64], 64, 3, 1, False),           # 24 % 16 != 0: falls back to the tap-streaming kernel
    (67, 16, [128], 128, 3, 1, True),         # many boxes, odd box count (MT = 2 tail)
    (3, 64, [128, 256], 128, 3, 1, True),     # transposed halo kernel (cout 128, 8 x 32 boxes): concat sources + residual
    (5, 32, [64], 384, 3, 1, False),          # transposed kernel, three 128-channel tiles per box, ragged batch
    (150, 32, [128], 128, 3, 1, True),        # transposed kernel: more tiles than SMs (persistent loop, both TMEM sets)
])
def test_igemm_conv_forward(N, H, chans, cout, ksize, stride, resid):
    err = _conv_case(N, H, chans, cout, ksize, stride, True, resid, seed=H + cout)
    assert err < 5e-3, err


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_igemm_tile_widths(bn):
    assert _conv_case(2, 16, [128], 256, 3, 1, True, False, seed=bn, bn=bn) < 5e-3


def test_igemm_plain_gemm_and_small_cout():
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(7)
    rows, K, Nn = 1000, 512, 384
    a = torch.randn(rows, K, generator=g).to(dev()).to(bf16)
    w = (torch.randn(Nn, K, generator=g) * K ** -0.5).to(dev()).to(bf16)
    bias = torch.randn(Nn, generator=g).to(dev())
    out = torch.empty(1, 1, rows, Nn, device=dev(), dtype=bf16)
    d = ops.make_igemm_desc([a.view(1, 1, rows, K)], [(0, 0, 0, 0, K // 64, 0)], w, out, Nn, bias=bias)
    ops.igemm(d)
    assert relerr(out.view(rows, Nn), a.float() @ w.float().t() + bias) < 5e-3
    # final eps conv: 128 -> 3 channels, NCHW fp32 output, weight rows padded to 16
    x = torch.randn(2, 128, 32, 32, generator=g).to(dev()).to(bf16).float()
    w3 = (torch.randn(3, 128, 3, 3, generator=g) * 0.03).to(dev()).to(bf16).float()
    b3 = torch.randn(3, generator=g).to(dev())
    wp = torch.zeros(16, 9 * 128, device=dev(), dtype=bf16)
    wp[:3] = pack_ohwi(w3)
    out3 = torch.empty(2, 3, 32, 32, device=dev())
    segs, _ = ops.conv_segments([128], 3)
    ops.igemm(ops.make_igemm_desc([nhwc(x)], segs, wp, out3, 3, bias=b3, out_mode=1))
    assert relerr(out3, F.conv2d(x, w3, b3, padding=1)) < 5e-3


def test_igemm_dgrad_matches_autograd():
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(8)
    N, H, cin, cout = 2, 16, 128, 192
    x = torch.randn(N, cin, H, H, generator=g).to(dev()).requires_grad_(True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.03).to(dev()).to(bf16).float()
    dy = torch.randn(N, cout, H, H, generator=g).to(dev()).to(bf16).float()
    F.conv2d(x, w, padding=1).backward(dy)
    wt = w.permute(1, 2, 3, 0).reshape(cin, 9 * cout).contiguous().to(bf16)     # [Cin][tap][Cout]
    segs, _ = ops.conv_segments([cout], 3, transposed=True)
    dx = torch.empty(N, H, H, cin, device=dev(), dtype=bf16)
    ops.igemm(ops.make_igemm_desc([nhwc(dy)], segs, wt, dx, cin))
    assert relerr(nchw(dx), x.grad) < 5e-3


@pytest.mark.parametrize("C0,C1,cout,H,film,skip", [(128, 0, 128, 64, False, 0), (128, 128, 128, 32, True, 0),
                                                     (256, 0, 256, 32, True, 128), (384, 256, 256, 32, False, 0),
                                                     # the pixel-major halo kernel: 16x16 level, N tiles 192 / 128 / 64
                                                     (384, 0, 384, 16, True, 0), (384, 384, 384, 16, False, 0),
                                                     (256, 0, 384, 16, True, 256), (128, 0, 128, 16, False, 0), (64, 64, 64, 48, True, 0)])
def test_conv_with_groupnorm_on_load_equals_the_two_pass_form(C0, C1, cout, H, film, skip):
    """inference: GroupNorm(+FiLM)+SiLU folded into the operand path of the 3x3 conv that consumes it
    (make_igemm_desc(gn=...), the transposed halo kernel's transform warps) == gn_apply_fwd followed by the conv, bit for
    bit; `skip`: a second, RAW source feeding a fused 1x1 skip conv (ResBlock conv2 with a channel change)"""
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(C0 + C1 + H)
    N, Ct = 3, C0 + C1
    x = (torch.randn(N, H, H, Ct, generator=g) * 1.3 + 0.2).to(dev()).to(bf16)
    x0 = x[..., :C0].contiguous()
    x1 = x[..., C0:].contiguous() if C1 else None
    xf = x.float()
    st = torch.stack([xf.sum(dim=(1, 2)), (xf * xf).sum(dim=(1, 2))], dim=-1)          # what the producing conv's epilogue leaves
    st0, st1 = st[:, :C0].contiguous(), (st[:, C0:].contiguous() if C1 else None)
    gamma = (1 + 0.2 * torch.randn(Ct, generator=g)).to(dev())
    beta = (0.2 * torch.randn(Ct, generator=g)).to(dev())
    film_t = (0.3 * torch.randn(N, 2 * Ct + 32, generator=g)).to(dev()) if film else None
    foff = 16
    xs = torch.randn(N, H, H, skip, generator=g).to(dev()).to(bf16) if skip else None
    segs, K = ops.conv_segments([C0, C1] if C1 else [C0], 3)
    Kt = K
    if skip:
        ssegs, Kt = ops.conv_segments([skip], 1, wk0=K, src0=2 if C1 else 1)
        segs = segs + ssegs
    w = (torch.randn(cout, Kt, generator=g) * Kt ** -0.5).to(dev()).to(bf16)
    bias = (0.1 * torch.randn(cout, generator=g)).to(dev())
    # two passes
    a, _, _ = ops.gn_apply_fwd(x0, st0, gamma, beta, x1=x1, stats1=st1, film=film_t, film_off=foff, silu=True)
    a0, a1 = a[..., :C0].contiguous(), (a[..., C0:].contiguous() if C1 else None)
    srcs_a = [a0] + ([a1] if C1 else []) + ([xs] if skip else [])
    stats_a = torch.zeros(N, cout, 2, device=dev())
    out_a = torch.empty(N, H, H, cout, device=dev(), dtype=bf16)
    ops.igemm(ops.make_igemm_desc(srcs_a, segs, w, out_a, cout, bias=bias, stats=stats_a))
    # fused
    ab = torch.full((N, Ct, 2), float("nan"), device=dev())
    ops.gn_apply_fwd(x0, st0, gamma, beta, x1=x1, stats1=st1, film=film_t, film_off=foff, silu=True, ab=ab, constants_only=True)
    srcs_x = [x0] + ([x1] if C1 else []) + ([xs] if skip else [])
    offs = [0] + ([C0] if C1 else []) + ([-1] if skip else [])
    stats_f = torch.zeros(N, cout, 2, device=dev())
    out_f = torch.full((N, H, H, cout), float("nan"), device=dev(), dtype=bf16)
    ops.igemm(ops.make_igemm_desc(srcs_x, segs, w, out_f, cout, bias=bias, stats=stats_f, gn=(ab, offs)))
    assert torch.isfinite(out_f.float()).all()
    assert torch.equal(out_f, out_a)
    assert torch.equal(stats_f, stats_a) or relerr(stats_f, stats_a) < 1e-6        # fp32 atomics: order may differ
    assert torch.equal(x0, x[..., :C0].contiguous())                                # the sources are not modified


@pytest.mark.parametrize("N,H,cin,cout,acc", [(2, 32, 128, 128, False), (3, 16, 256, 192, True), (2, 8, 384, 384, False)])
def test_stride2_dgrad_parity_classes_match_autograd(N, H, cin, cout, acc):
    """data gradient of a 3x3 stride-2 pad-1 conv (Downsample, ref unet.py:97-105) as four launches over the low-resolution
    dy, one per output parity, stored with pixel stride 2 (cdae_igemm_desc.sps / ooh / oow); acc: added to what dx holds"""
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(H + cin)
    x = torch.randn(N, cin, H, H, generator=g).to(dev()).requires_grad_(True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.03).to(dev()).to(bf16).float()
    dy = torch.randn(N, cout, H // 2, H // 2, generator=g).to(dev()).to(bf16).float()
    F.conv2d(x, w, stride=2, padding=1).backward(dy)
    wt = w.permute(1, 2, 3, 0).reshape(cin, 9 * cout).contiguous().to(bf16)     # [Cin][tap][Cout]
    base = torch.randn(N, H, H, cin, generator=g).to(dev()).to(bf16)
    dx = base.clone() if acc else torch.full((N, H, H, cin), float("nan"), device=dev(), dtype=bf16)
    for a in range(2):
        for b in range(2):
            khs = [(1, 0)] if a == 0 else [(0, 1), (2, 0)]
            kws = [(1, 0)] if b == 0 else [(0, 1), (2, 0)]
            segs = [(0, dh, dw, 0, cout // 64, (kh * 3 + kw) * cout) for kh, dh in khs for kw, dw in kws]
            ops.igemm(ops.make_igemm_desc([nhwc(dy)], segs, wt, dx, cin, resid=dx if acc else None, sps=2, ooh=a, oow=b))
    ref = x.grad + (nchw(base) if acc else 0)
    assert torch.isfinite(dx.float()).all()
    assert relerr(nchw(dx), ref) < 5e-3


@pytest.mark.parametrize("N,H,cin,cout,ksize,stride", [(2, 32, 128, 128, 3, 1), (4, 8, 256, 128, 3, 1), (2, 16, 64, 64, 3, 1),
                                                       (3, 16, 192, 256, 1, 1), (2, 32, 128, 128, 3, 2), (8, 4, 128, 128, 3, 1),
                                                       (2, 64, 64, 128, 3, 1), (3, 12, 128, 192, 3, 1), (2, 16, 256, 256, 3, 1), (5, 8, 512, 64, 1, 1)])
def test_wgrad_matches_autograd(N, H, cin, cout, ksize, stride):
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(H + cin)
    x = torch.randn(N, cin, H, H, generator=g).to(dev()).to(bf16).float()
    w = torch.zeros(cout, cin, ksize, ksize, device=dev(), requires_grad=True)
    y = F.conv2d(x, w, stride=stride, padding=ksize // 2)
    dy = torch.randn(y.shape, generator=g).to(dev()).to(bf16).float()
    y.backward(dy)
    dw = torch.zeros(cout, ksize * ksize, cin, device=dev())
    fused_bias = True                                # every weight-gradient kernel also produces the bias gradient
    db = torch.full((cout,), 0.5, device=dev()) if fused_bias else None
    ops.wgrad(ops.make_wgrad_desc(nhwc(dy), nhwc(x), dw, cout, cin, ksize=ksize, in_stride=stride, dbias=db))
    ref = w.grad.permute(0, 2, 3, 1).reshape(cout, ksize * ksize, cin)
    assert relerr(dw, ref) < 5e-3
    if fused_bias:
        assert relerr(db - 0.5, dy.sum(dim=(0, 2, 3))) < 2e-3
    # accumulation semantics (+=) and a channel window of a wider source (concat operand)
    if cin >= 128 and ksize == 3 and stride == 1:
        dw2 = torch.ones(cout, 9, cin, device=dev())
        ops.wgrad(ops.make_wgrad_desc(nhwc(dy), nhwc(x), dw2, cout, 64, ksize=3, c0=64, ci_off=64, dw_ld=cin))
        assert relerr(dw2[:, :, 64:128] - 1, ref[:, :, 64:128]) < 5e-3
        assert float((dw2[:, :, :64] - 1).abs().max()) == 0


# ------------------------------------------------------------------ attention
# ------------------------------------------------------------------ GroupNorm statistics from the conv epilogue + streaming GN
@pytest.mark.parametrize("N,H,chans,cout,ksize,stride,resid", [
    (2, 64, [64], 128, 3, 1, False),          # halo kernel, N tile 128 (two slabs per tile)
    (3, 32, [128], 256, 3, 1, True),          # residual epilogue, N tile 256
    (5, 8, [256], 256, 3, 1, False),          # two images per 128-pixel box, ragged batch
    (2, 32, [128], 128, 3, 2, False),         # stride-2 (Downsample)
    (2, 16, [192], 64, 1, 1, True),           # 1x1 conv, cout 64
    (2, 48, [64, 128], 192, 3, 1, True),      # non-power-of-two image: partial boxes must be masked out of the sums
    (3, 24, [64], 64, 3, 1, False),           # tap-streaming kernel with partial boxes
    (67, 16, [128], 128, 3, 1, True),         # many boxes (MT = 2 tail)
    (3, 32, [128, 64], 128, 3, 1, True),      # transposed halo kernel: per-thread channel sums
    (70, 64, [64], 128, 3, 1, False),         # transposed kernel, persistent loop over > 148 tiles
])
def test_igemm_channel_stats(N, H, chans, cout, ksize, stride, resid):
    """cdae_igemm_desc.stats: per-(image, channel) sum / sum of squares of the bf16 output, accumulated by the epilogue"""
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(H * 7 + cout)
    cin = sum(chans)
    x = torch.randn(N, cin, H, H, generator=g).to(dev()).to(bf16)
    w = (torch.randn(cout, cin, ksize, ksize, generator=g) * (cin * ksize * ksize) ** -0.5).to(dev()).to(bf16).float()
    b = torch.randn(cout, generator=g).to(dev())
    OH = (H + stride - 1) // stride
    r = torch.randn(N, OH, OH, cout, generator=g).to(dev()).to(bf16) if resid else None
    xn = x.permute(0, 2, 3, 1).contiguous()
    xs, off = [], 0
    for c in chans:
        xs.append(xn[..., off:off + c].contiguous()); off += c
    segs, _ = ops.conv_segments(chans, ksize)
    out = torch.full((N, OH, OH, cout), float("nan"), device=dev(), dtype=bf16)
    stats = torch.zeros(N, cout, 2, device=dev())
    d = ops.make_igemm_desc(xs, segs, pack_ohwi(w), out, cout, in_stride=stride, bias=b, resid=r, stats=stats)
    ops.igemm(d)
    o = out.double()
    ref = torch.stack([o.sum(dim=(1, 2)), (o * o).sum(dim=(1, 2))], dim=-1)
    assert bool(torch.isfinite(stats).all())
    np.testing.assert_allclose(stats.double().cpu().numpy(), ref.cpu().numpy(), rtol=2e-4, atol=2e-3 * OH)
    ops.igemm(d)    # accumulates
    np.testing.assert_allclose(stats.double().cpu().numpy(), 2 * ref.cpu().numpy(), rtol=2e-4, atol=4e-3 * OH)


def test_igemm_channel_stats_rejects_unsupported():
    from causaldiffae_b200 import ops
    from causaldiffae_b200._lib import CdaeError
    x = torch.randn(2, 4, 4, 64, device=dev()).to(bf16)          # 16 pixels per image: a warp's rows span two images
    w = torch.randn(64, 9 * 64, device=dev()).to(bf16)
    segs, _ = ops.conv_segments([64], 3)
    d = ops.make_igemm_desc([x], segs, w, torch.empty(2, 4, 4, 64, device=dev(), dtype=bf16), 64,
                            stats=torch.zeros(2, 64, 2, device=dev()))
    with pytest.raises(CdaeError):
        ops.igemm(d)


@pytest.mark.parametrize("N,H,C,cout", [(3, 16, 128, 128), (2, 32, 256, 256), (5, 16, 384, 384), (2, 16, 64, 128)])
def test_conv_over_nearest_upsampling_on_load_equals_materialised(N, H, C, cout):
    """Upsample (ref unet.py:69-79): conv3x3(F.interpolate(x, scale_factor=2, mode="nearest")) with the expansion done by the
    conv's transform warps from a low-resolution halo tile (make_igemm_desc(up2x=True)) == upsample2x + conv, bit for bit,
    and both against torch"""
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(N + H + C)
    x = torch.randn(N, H, H, C, generator=g).to(dev()).to(bf16)
    segs, K = ops.conv_segments([C], 3)
    wt = (torch.randn(cout, C, 3, 3, generator=g) * (9 * C) ** -0.5).to(dev()).to(bf16).float()
    w = pack_ohwi(wt)
    bias = (0.1 * torch.randn(cout, generator=g)).to(dev())
    u = ops.upsample2x(x)
    out_a = torch.empty(N, 2 * H, 2 * H, cout, device=dev(), dtype=bf16)
    st_a = torch.zeros(N, cout, 2, device=dev())
    ops.igemm(ops.make_igemm_desc([u], segs, w, out_a, cout, bias=bias, stats=st_a))
    out_f = torch.full((N, 2 * H, 2 * H, cout), float("nan"), device=dev(), dtype=bf16)
    st_f = torch.zeros(N, cout, 2, device=dev())
    ops.igemm(ops.make_igemm_desc([x], segs, w, out_f, cout, bias=bias, stats=st_f, up2x=True))
    assert torch.equal(out_f, out_a)
    assert relerr(st_f, st_a) < 1e-6
    ref = F.conv2d(F.interpolate(nchw(x), scale_factor=2, mode="nearest"), wt, bias, padding=1)
    assert relerr(nchw(out_f), ref) < 5e-3


def test_head_conv_with_groupnorm_on_load_fp32_nchw_output():
    """the UNet head (GroupNorm -> SiLU -> conv 128 -> 3, fp32 NCHW output, ref unet.py:498) with the norm applied on load"""
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(11)
    N, H, C = 3, 32, 128
    x = (torch.randn(N, H, H, C, generator=g) * 1.2 - 0.1).to(dev()).to(bf16)
    xf = x.float()
    st = torch.stack([xf.sum(dim=(1, 2)), (xf * xf).sum(dim=(1, 2))], dim=-1).contiguous()
    gamma, beta = (1 + 0.2 * torch.randn(C, generator=g)).to(dev()), (0.2 * torch.randn(C, generator=g)).to(dev())
    wp = torch.zeros(16, 9 * C, device=dev(), dtype=bf16)
    wp[:3] = (torch.randn(3, 9 * C, generator=g) * 0.03).to(dev()).to(bf16)
    b3 = torch.randn(3, generator=g).to(dev())
    segs, _ = ops.conv_segments([C], 3)
    a, _, _ = ops.gn_apply_fwd(x, st, gamma, beta, silu=True)
    out_a = torch.empty(N, 3, H, H, device=dev())
    ops.igemm(ops.make_igemm_desc([a], segs, wp, out_a, 3, bias=b3, out_mode=1))
    ab = torch.empty(N, C, 2, device=dev())
    ops.gn_apply_fwd(x, st, gamma, beta, silu=True, ab=ab, constants_only=True)
    out_f = torch.full((N, 3, H, H), float("nan"), device=dev())
    ops.igemm(ops.make_igemm_desc([x], segs, wp, out_f, 3, bias=b3, out_mode=1, gn=(ab, [0])))
    assert torch.equal(out_f, out_a)


def test_igemm_rejects_unsupported_groupnorm_on_load_and_output_placement():
    """the new descriptor fields fail loudly where the kernels cannot honour them (no silent fallback)"""
    from causaldiffae_b200 import ops
    from causaldiffae_b200._lib import CdaeError
    segs, _ = ops.conv_segments([64], 3)
    w = torch.randn(128, 9 * 64, device=dev()).to(bf16)
    # GroupNorm on load on an 8x8 image: does not tile into 8 x 16 boxes (no halo kernel)
    x = torch.randn(2, 8, 8, 64, device=dev()).to(bf16)
    ab = torch.zeros(2, 64, 2, device=dev())
    d = ops.make_igemm_desc([x], segs, w, torch.empty(2, 8, 8, 128, device=dev(), dtype=bf16), 128, gn=(ab, [0]))
    with pytest.raises(CdaeError):
        ops.igemm(d)
    # ... and a table that is too narrow for the source
    x = torch.randn(2, 32, 32, 64, device=dev()).to(bf16)
    d = ops.make_igemm_desc([x], segs, w, torch.empty(2, 32, 32, 128, device=dev(), dtype=bf16), 128,
                            gn=(torch.zeros(2, 32, 2, device=dev()), [0]))
    with pytest.raises(CdaeError):
        ops.igemm(d)
    # strided output placement together with channel statistics
    dy = torch.randn(2, 16, 16, 64, device=dev()).to(bf16)
    d = ops.make_igemm_desc([dy], segs, w, torch.empty(2, 32, 32, 128, device=dev(), dtype=bf16), 128,
                            stats=torch.zeros(2, 128, 2, device=dev()), sps=2, ooh=1, oow=0)
    with pytest.raises(CdaeError):
        ops.igemm(d)
    # an offset outside the 2 x 2 parity classes
    d = ops.make_igemm_desc([dy], segs, w, torch.empty(2, 32, 32, 128, device=dev(), dtype=bf16), 128, sps=2, ooh=2, oow=0)
    with pytest.raises(CdaeError):
        ops.igemm(d)


@pytest.mark.parametrize("N,H,C0,C1,cin,ksize,film,silu", [
    (2, 64, 128, 0, 128, 3, True, True),        # halo kernel, one image per box, FiLM (ResBlock out_layers norm)
    (3, 32, 256, 128, 256, 3, False, True),     # concat input of the norm: x slab comes from two tensors
    (5, 8, 512, 512, 512, 3, False, True),      # two images per 128-pixel box (tap-streaming kernel), ragged batch
    (2, 16, 384, 0, 1152, 1, False, False),     # attention norm: 1x1 qkv data gradient, no SiLU
    (2, 48, 64, 64, 64, 3, True, True),         # non-power-of-two image: partial boxes masked out of the sums
    (67, 16, 128, 0, 128, 3, True, True),       # many boxes (MT = 2 tail)
    (3, 32, 64, 64, 256, 3, True, True),        # transposed halo kernel (128 output channels of the data gradient), concat x
    (40, 64, 128, 0, 128, 3, False, False),     # transposed kernel, no SiLU, persistent loop
])
def test_groupnorm_backward_from_dgrad_epilogue(N, H, C0, C1, cin, ksize, film, silu):
    """The fused GroupNorm backward: y = act(FiLM(GN(x))) feeds a conv; that conv's DATA-GRADIENT launch (gnb_*) stores
    du = dy * silu'(u) and accumulates {sum du, sum du*x} per (image, channel); cdae_gn_bwd_apply is then one streaming
    pass.  Checked against fp32 torch autograd of the same composition (GroupNorm -> FiLM -> SiLU -> conv)."""
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(N * 131 + H + C0)
    Ct = C0 + C1
    cout_conv = cin                                   # the consumer conv: Ct -> cout_conv channels
    x = (torch.randn(N, Ct, H, H, generator=g) * 1.5 + 0.3).to(dev()).to(bf16).float()
    gamma = (1 + 0.2 * torch.randn(Ct, generator=g)).to(dev())
    beta = (0.2 * torch.randn(Ct, generator=g)).to(dev())
    film_t = (0.3 * torch.randn(N, 2 * Ct + 32, generator=g)).to(dev()) if film else None
    foff = 16
    w = (torch.randn(cout_conv, Ct, ksize, ksize, generator=g) * (Ct * ksize * ksize) ** -0.5).to(dev()).to(bf16).float()
    dz = torch.randn(N, cout_conv, H, H, generator=g).to(dev()).to(bf16).float()      # gradient w.r.t. the conv output
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    fr = film_t.clone().requires_grad_(True) if film else None
    u = F.group_norm(xr, 32, gr, br, eps=1e-5)
    if film:
        u = u * (1 + fr[:, foff:foff + Ct, None, None]) + fr[:, foff + Ct:foff + 2 * Ct, None, None]
    a = u * torch.sigmoid(u) if silu else u
    F.conv2d(a, w, padding=ksize // 2).backward(dz)
    # CUDA path: streaming forward (leaves the {a, b} table), data-gradient conv with the fused statistics, streaming apply
    x_nhwc = nhwc(x)
    x0 = x_nhwc[..., :C0].contiguous()
    x1 = x_nhwc[..., C0:].contiguous() if C1 else None
    st = torch.stack([x.sum(dim=(2, 3)), (x * x).sum(dim=(2, 3))], dim=-1)
    ab = torch.full((N, Ct, 2), float("nan"), device=dev())
    y, mean, rstd = ops.gn_apply_fwd(x0, st[:, :C0].contiguous(), gamma, beta, x1=x1, stats1=st[:, C0:].contiguous() if C1 else None,
                                     film=film_t, film_off=foff, silu=silu, ab=ab)
    assert relerr(nchw(y), a) < 6e-3
    wt = w.permute(1, 2, 3, 0).reshape(Ct, -1).contiguous().to(bf16)       # [Ct, taps*cout]: transposed packing of the engine
    segs, _ = ops.conv_segments([cout_conv], ksize, transposed=True)
    du = torch.full((N, H, H, Ct), float("nan"), device=dev(), dtype=bf16)
    ws = torch.zeros(N, Ct, 2, device=dev())
    d = ops.make_igemm_desc([nhwc(dz)], segs, wt, du, Ct, gnb=dict(x0=x0, x1=x1, ab=ab if silu else None, ws=ws, silu=silu))
    ops.igemm(d)
    dgamma, dbeta = torch.zeros(Ct, device=dev()), torch.zeros(Ct, device=dev())
    dfilm = torch.zeros_like(film_t) if film else None
    dx0, dx1 = ops.gn_bwd_apply(du, x0, gamma, beta, mean, rstd, ws, x1=x1, film=film_t, film_off=foff, dgamma=dgamma,
                                dbeta=dbeta, dfilm=dfilm)
    dx = torch.cat([dx0, dx1], dim=-1) if C1 else dx0
    # the statistics are those of du as stored
    duf, xf = du.double(), x_nhwc.double()
    ref_ws = torch.stack([duf.sum(dim=(1, 2)), (duf * xf).sum(dim=(1, 2))], dim=-1)
    np.testing.assert_allclose(ws.double().cpu().numpy(), ref_ws.cpu().numpy(), rtol=3e-4, atol=3e-3 * H)
    assert relerr(nchw(dx), xr.grad) < 1e-2, relerr(nchw(dx), xr.grad)
    assert relerr(dgamma, gr.grad) < 5e-3 and relerr(dbeta, br.grad) < 5e-3
    if film:
        assert relerr(dfilm, fr.grad) < 5e-3
    # accumulate / dadd modes of the apply pass
    base = torch.randn_like(dx0.float()).to(bf16)
    acc0, acc1 = base.clone(), (torch.zeros_like(dx1) if C1 else None)
    dadd = torch.randn(N, H, H, Ct, generator=g).to(dev()).to(bf16)
    ops.gn_bwd_apply(du, x0, gamma, beta, mean, rstd, ws, x1=x1, film=film_t, film_off=foff, dx0=acc0, dx1=acc1,
                     accumulate_dx=True, dadd=dadd)
    assert relerr(acc0.float(), base.float() + dx0.float() + dadd[..., :C0].float()) < 1e-2
    if C1:
        assert relerr(acc1.float(), dx1.float() + dadd[..., C0:].float()) < 1e-2


@pytest.mark.parametrize("C0,C1,HW,film,silu,B", [(128, 0, 64 * 64, True, True, 3), (64, 0, 16 * 16, False, True, 2),
                                                  (512, 384, 8 * 8, False, True, 3), (256, 128, 32 * 32, True, True, 2),
                                                  (384, 0, 16 * 16, False, False, 5), (384, 256, 16 * 16, True, True, 2),
                                                  (512, 512, 8 * 8, False, True, 70), (128, 128, 48 * 48, False, True, 2)])
def test_groupnorm_apply_fwd_matches_reducing_kernel_and_torch(C0, C1, HW, film, silu, B):
    from causaldiffae_b200 import ops
    g = torch.Generator().manual_seed(C0 + 3 * C1 + HW)
    Ct = C0 + C1
    S = int(HW ** 0.5)
    x = (torch.randn(B, Ct, S, S, generator=g) * 1.5 + 0.3).to(dev()).to(bf16).float()
    gamma = (1 + 0.2 * torch.randn(Ct, generator=g)).to(dev())
    beta = (0.2 * torch.randn(Ct, generator=g)).to(dev())
    film_t = (0.3 * torch.randn(B, 2 * Ct + 32, generator=g)).to(dev()) if film else None
    foff = 16
    u = F.group_norm(x, 32, gamma, beta, eps=1e-5)
    if film:
        u = u * (1 + film_t[:, foff:foff + Ct, None, None]) + film_t[:, foff + Ct:foff + 2 * Ct, None, None]
    yref = u * torch.sigmoid(u) if silu else u
    x_nhwc = nhwc(x)
    x0 = x_nhwc[..., :C0].contiguous()
    x1 = x_nhwc[..., C0:].contiguous() if C1 else None
    st = torch.stack([x.sum(dim=(2, 3)), (x * x).sum(dim=(2, 3))], dim=-1)          # [B, Ct, 2]
    st0 = st[:, :C0].contiguous()
    st1 = st[:, C0:].contiguous() if C1 else None
    y, mean, rstd = ops.gn_apply_fwd(x0, st0, gamma, beta, x1=x1, stats1=st1, film=film_t, film_off=foff, silu=silu)
    assert relerr(nchw(y), yref) < 6e-3
    y2, mean2, rstd2 = ops.gn_fwd(x0, gamma, beta, x1=x1, film=film_t, film_off=foff, silu=silu)
    np.testing.assert_allclose(mean.cpu().numpy(), mean2.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rstd.cpu().numpy(), rstd2.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert relerr(y.float(), y2.float()) < 4e-3


@pytest.mark.parametrize("B,T,heads,ch", [(2, 256, 4, 96), (3, 64, 4, 128), (2, 16, 4, 32), (1, 784, 2, 64), (2, 100, 1, 48)])
def test_attention_fwd_bwd(B, T, heads, ch):
    from causaldiffae_b200 import ops
    from oracle.model import qkv_attention
    g = torch.Generator().manual_seed(T + ch)
    C = heads * ch
    qkv = torch.randn(B, T, 3 * C, generator=g).to(dev()).to(bf16)
    dout = torch.randn(B, T, C, generator=g).to(dev()).to(bf16)
    # oracle layout: [B*heads, 3*ch, T]
    ref_in = qkv.float().permute(0, 2, 1).reshape(B * heads, 3 * ch, T).clone().requires_grad_(True)
    ref = qkv_attention(ref_in)                                        # [B*heads, ch, T]
    ref.backward(dout.float().permute(0, 2, 1).reshape(B * heads, ch, T))
    out, lse = ops.attn_fwd(qkv, heads)
    assert relerr(out.float().permute(0, 2, 1).reshape(B * heads, ch, T), ref) < 1e-2
    dqkv = ops.attn_bwd(qkv, out, dout, lse, heads)
    got = dqkv.float().permute(0, 2, 1).reshape(B * heads, 3 * ch, T)
    for name, sl in (("dq", slice(0, ch)), ("dk", slice(ch, 2 * ch)), ("dv", slice(2 * ch, 3 * ch))):
        assert relerr(got[:, sl], ref_in.grad[:, sl]) < 2e-2, name


# ------------------------------------------------------------------ causal DAG mask layer (fused) vs the two reference methods
@pytest.mark.parametrize("B,n,latent", [(64, 4, 512), (5, 2, 512), (33, 4, 256)])
def test_dag_layer_fwd_bwd(B, n, latent):
    from causaldiffae_b200.nn import CausalModeling
    torch.manual_seed(B + n)
    mod = CausalModeling(latent_dim=latent, num_var=n).to(dev())
    A = torch.triu(torch.ones(n, n), diagonal=1).to(dev())
    A[0, -1] = 0.0
    u = torch.randn(B, latent, device=dev(), requires_grad=True)
    dz = torch.randn(B, latent, device=dev())
    # reference composition (ref nn.py:290-312) through plain torch autograd
    ref = mod.nonlinearity_add_back_noise(u, mod.causal_masking(u, A))
    ref.backward(dz)
    gref = [p.grad.clone() for p in mod._mlp_params()]
    duref = u.grad.clone()
    for p in mod._mlp_params():
        p.grad = torch.full_like(p, 0.25)            # the kernel ACCUMULATES into existing gradients
    u.grad = None
    assert mod.fused_ok(u)
    out = mod(u, A)
    assert relerr(out, ref.detach()) < 1e-5
    out.backward(dz)
    assert relerr(u.grad, duref) < 1e-5
    for p, g in zip(mod._mlp_params(), gref):
        assert relerr(p.grad - 0.25, g) < 1e-4
    # workspace is left zeroed for the next call
    assert float(mod._workspace(u).abs().max()) == 0.0


# ------------------------------------------------------------------ empty inputs
def test_empty_batch_is_a_no_op_everywhere():
    """a rank whose shard is empty (ragged sharding of interventions, shard_range with n < world) must be able to call
    the whole path: every entry point accepts a zero batch, launches nothing and returns correctly shaped tensors"""
    from causaldiffae_b200 import ops
    d = dev()
    z4 = torch.zeros(0, 3, 8, 8, device=d)
    t0 = torch.zeros(0, dtype=torch.int64, device=d)
    tab = torch.rand(10, device=d)
    assert ops.q_sample(z4, z4, t0, tab, tab).shape == (0, 3, 8, 8)
    mse, dp = ops.mse_loss(z4, z4, torch.zeros(0, device=d), want_grad=True)
    assert mse.shape == (0,) and dp.shape == z4.shape
    coef = torch.rand(5, 8, device=d)
    out, x0 = ops.ddim_step(z4, z4, coef, torch.zeros(1, dtype=torch.int32, device=d), want_xstart=True)
    assert out.shape == z4.shape and x0.shape == z4.shape
    xb = torch.zeros(0, 8, 8, 64, device=d, dtype=bf16)
    g, b = torch.ones(64, device=d), torch.zeros(64, device=d)
    y, mean, rstd = ops.gn_fwd(xb, g, b)
    assert y.shape == xb.shape and mean.shape == (0, 32)
    y2, _, _ = ops.gn_apply_fwd(xb, torch.zeros(0, 64, 2, device=d), g, b)
    assert y2.shape == xb.shape
    dx0, _ = ops.gn_bwd(xb, xb, g, b, mean, rstd)
    assert dx0.shape == xb.shape
    w = torch.zeros(64, 9 * 64, device=d, dtype=bf16)
    segs, _ = ops.conv_segments([64], 3)
    o = torch.empty(0, 8, 8, 64, device=d, dtype=bf16)
    ops.igemm(ops.make_igemm_desc([xb], segs, w, o, 64, stats=torch.zeros(0, 64, 2, device=d)))
    dw = torch.zeros(64, 9, 64, device=d)
    ops.wgrad(ops.make_wgrad_desc(o, xb, dw, 64, 64))
    assert float(dw.abs().max()) == 0.0
    qkv = torch.zeros(0, 16, 3 * 64, device=d, dtype=bf16)
    ao, lse = ops.attn_fwd(qkv, 4)
    assert ao.shape == (0, 16, 64)
    im = torch.zeros(4, 8, 8, 3, dtype=torch.uint8, device=d)
    gx, gl = ops.gather_images(im, t0, labels=torch.zeros(4, 2, device=d))
    assert gx.shape == (0, 3, 8, 8) and gl.shape == (0, 2)
    torch.cuda.synchronize()

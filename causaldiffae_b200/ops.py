"""Tensor-level wrappers over the C ABI (one function per libcdae entry point).

These are thin: they check dtype/layout, allocate outputs with torch (device memory is PyTorch's job) and launch on
the current stream.  Everything numerical happens in the hand-written kernels; there is no torch fallback.
Activations are NHWC bf16 tensors [N, H, W, C] (or [rows, C] for plain GEMMs)."""
import ctypes as C

import torch

from . import _lib
from ._lib import IgemmDesc, WgradDesc, SgemmDesc, Seg, ptr, stream, check


# ---------------------------------------------------------------------------------------------------- side stream
# Launch sequences fork independent work (weight gradients, the bf16 weight pack) onto ONE extra stream per device; under
# stream capture the wait_stream pairs become graph edges, so the branches run concurrently inside the replayed graph.
_SIDE = {}


def side_stream(which=0):
    key = (torch.cuda.current_device(), which)
    s = _SIDE.get(key)
    if s is None:
        s = _SIDE[key] = torch.cuda.Stream()
    return s


class on_side:
    """with on_side(): launches inside run on the side stream, ordered after everything issued so far on the current one;
    join_side() orders the current stream after them.  `which` selects one of several independent side streams."""

    def __init__(self, which=0):
        self.which = which

    def __enter__(self):
        self.side = side_stream(self.which)
        self.side.wait_stream(torch.cuda.current_stream())
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self.ctx.__exit__(*exc)


def join_side(which=0):
    torch.cuda.current_stream().wait_stream(side_stream(which))

bf16 = torch.bfloat16


def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expected contiguous fp32 CUDA tensor"
    return t


def _bf16c(t):
    assert t.is_cuda and t.dtype == bf16 and t.is_contiguous(), "expected contiguous bf16 CUDA tensor"
    return t


# ------------------------------------------------------------------ diffusion elementwise
def q_sample(x0, noise, t, sqrt_ac, sqrt_1mac, out=None):
    """ref gaussian_diffusion.py:201-222"""
    _f32c(x0); _f32c(noise); _f32c(sqrt_ac); _f32c(sqrt_1mac)
    assert t.dtype == torch.int64 and t.is_cuda and x0.shape == noise.shape and t.numel() == x0.shape[0]
    t = t.contiguous()          # an expanded (stride-0) view would be read past its one element
    out = torch.empty_like(x0) if out is None else out
    B = x0.shape[0]
    if x0.numel() == 0:
        return out
    check(_lib.lib().cdae_q_sample(ptr(x0), ptr(noise), ptr(t), ptr(sqrt_ac), ptr(sqrt_1mac), ptr(out), B,
                                   x0.numel() // max(B, 1), stream()))
    return out


def mse_loss(pred, target, gscale=None, want_grad=False, gmul=1.0, mse=None, dpred=None):
    """per-sample mean((target-pred)^2) and, optionally, d(sum_b gmul*gscale[b]*mse[b])/dpred (ref gaussian_diffusion.py:847)."""
    _f32c(pred); _f32c(target)
    B = pred.shape[0]
    mse = torch.empty(B, device=pred.device, dtype=torch.float32) if mse is None else mse
    if want_grad:
        dpred = torch.empty_like(pred) if dpred is None else dpred
        _f32c(gscale)
    check(_lib.lib().cdae_mse_loss(ptr(pred), ptr(target), ptr(mse), ptr(gscale), float(gmul), ptr(dpred), B,
                                   pred.numel() // max(B, 1), stream()))
    return mse, dpred


def ddim_step(x, eps_c, coef_table, t_idx, eps_u=None, w=None, noise=None, want_xstart=False, out=None):
    """ref gaussian_diffusion.py:277-285, 320-341, 506-558 (coef_table rows built by SpacedDiffusion.ddim_coef_table)."""
    _f32c(x); _f32c(eps_c); _f32c(coef_table)
    assert t_idx.dtype == torch.int32 and t_idx.is_cuda
    B = x.shape[0]
    stride = 0 if t_idx.numel() == 1 else 1
    assert stride == 0 or t_idx.numel() == B
    out = torch.empty_like(x) if out is None else out
    x0 = torch.empty_like(x) if want_xstart else None
    check(_lib.lib().cdae_ddim_step(ptr(x), ptr(eps_c), ptr(eps_u), float(w if w is not None else 0.0),
                                    int(w is not None), ptr(coef_table), ptr(t_idx), stride, ptr(noise), ptr(out),
                                    ptr(x0), B, x.numel() // max(B, 1), stream()))
    return out, x0


def adam_ema(p, g, m, v, ema, hyper, step, gsq_out=None, guard=None, lognorm=None):
    """ref train_util.py:276-303 + nn.py:503-513; flat fp32 arenas, hyper = device fp32[7] {lr, b1, b2, eps, wd, ema_rate,
    grad_scale}, step = device int64[1] (incremented by the launch), g fp32 or bf16, guard = device fp32[1] or None."""
    n = p.numel()
    assert step.dtype == torch.int64 and g.numel() == n
    check(_lib.lib().cdae_adam_ema(ptr(p), ptr(g), int(g.dtype == bf16), ptr(m), ptr(v), ptr(ema), ptr(hyper), ptr(step),
                                   ptr(guard), ptr(gsq_out), ptr(lognorm), n, stream()))


def sumsq(g, out):
    """out (fp32[1]) += sum(g^2); g fp32 or bf16"""
    assert g.is_cuda and g.is_contiguous() and g.dtype in (torch.float32, bf16)
    check(_lib.lib().cdae_sumsq(ptr(g), int(g.dtype == bf16), ptr(out), g.numel(), stream()))
    return out


def cast_bf16(src, out=None):
    out = torch.empty(src.shape, device=src.device, dtype=bf16) if out is None else out
    check(_lib.lib().cdae_cast_bf16(ptr(_f32c(src)), ptr(out), src.numel(), stream()))
    return out


def ema_update(ema, p, rate):
    check(_lib.lib().cdae_ema_update(ptr(ema), ptr(p), float(rate), p.numel(), stream()))


def zero_(t):
    check(_lib.lib().cdae_zero(ptr(t), t.numel() * t.element_size(), stream()))
    return t


# ------------------------------------------------------------------ layout
def nchw_to_nhwc_pad(x, cpad=64, out=None):
    _f32c(x)
    N, Cc, H, W = x.shape
    out = torch.empty(N, H, W, cpad, device=x.device, dtype=bf16) if out is None else out
    check(_lib.lib().cdae_nchw_to_nhwc_pad(ptr(x), ptr(out), N, Cc, H, W, cpad, stream()))
    return out


def nhwc_to_nchw(x, c):
    _bf16c(x)
    N, H, W, ld = x.shape
    out = torch.empty(N, c, H, W, device=x.device, dtype=torch.float32)
    check(_lib.lib().cdae_nhwc_to_nchw(ptr(x), ptr(out), N, c, H, W, ld, stream()))
    return out


def upsample2x(x, out=None):
    _bf16c(x)
    N, H, W, Cc = x.shape
    out = torch.empty(N, 2 * H, 2 * W, Cc, device=x.device, dtype=bf16) if out is None else out
    check(_lib.lib().cdae_upsample2x(ptr(x), ptr(out), N, H, W, Cc, stream()))
    return out


def sumpool2x(dy, out=None, accumulate=False):
    _bf16c(dy)
    N, H2, W2, Cc = dy.shape
    out = torch.empty(N, H2 // 2, W2 // 2, Cc, device=dy.device, dtype=bf16) if out is None else out
    check(_lib.lib().cdae_sumpool2x(ptr(dy), ptr(out), N, H2 // 2, W2 // 2, Cc, int(accumulate), stream()))
    return out


def zero_insert2x(x, out=None):
    _bf16c(x)
    N, H, W, Cc = x.shape
    out = torch.empty(N, 2 * H, 2 * W, Cc, device=x.device, dtype=bf16) if out is None else out
    check(_lib.lib().cdae_zero_insert2x(ptr(x), ptr(out), N, H, W, Cc, stream()))
    return out


def colsum_(x2d, out, c=None, groups=1, out_ld=0):
    """out[c] += sum_rows x2d[:, c]  (bias gradients); groups > 1: x2d is `groups` consecutive row blocks and
    out[g * out_ld + c] gets the sums of block g (per-image sums)."""
    _bf16c(x2d)
    rows, ld = x2d.reshape(-1, x2d.shape[-1]).shape
    assert rows % groups == 0
    check(_lib.lib().cdae_colsum(ptr(x2d), ptr(out), rows // groups, c if c is not None else ld, ld, groups, out_ld, stream()))
    return out


def dropout_(x, state, layer_offset, p_dev):
    """inverted dropout in place (bf16); the same call on the gradient is its backward.  state: device int64[2] {seed, base},
    p_dev: device fp32[1] (0 = identity)"""
    _bf16c(x)
    check(_lib.lib().cdae_dropout(ptr(x), x.numel(), ptr(state), int(layer_offset), ptr(p_dev), stream()))
    return x


def gather_images(images_u8, idx, labels=None, mode=0, out=None, out_labels=None):
    """batch assembly from an HBM-resident dataset (ref image_datasets.py __getitem__ + DataLoader collate):
    images_u8 uint8 [n,H,W,C], idx int64 [B], labels fp32 [n,L] -> fp32 [B,C,H,W] = u8/255 (mode 1: u8/127.5-1), [B,L]"""
    assert images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.is_contiguous() and images_u8.dim() == 4
    assert idx.is_cuda and idx.dtype == torch.int64 and idx.is_contiguous()
    n, H, W, Cc = images_u8.shape
    B = idx.numel()
    L = 0
    if labels is not None:
        _f32c(labels); L = labels.shape[1]
        out_labels = torch.empty(B, L, device=idx.device, dtype=torch.float32) if out_labels is None else out_labels
    out = torch.empty(B, Cc, H, W, device=idx.device, dtype=torch.float32) if out is None else out
    check(_lib.lib().cdae_gather_images(ptr(images_u8), ptr(labels), ptr(idx), ptr(out), ptr(out_labels), B, H, W, Cc, L,
                                        int(mode), stream()))
    return out, out_labels


# ------------------------------------------------------------------ GroupNorm32 (+FiLM)(+SiLU)
def gn_fwd(x0, gamma, beta, x1=None, film=None, film_off=0, silu=True, out=None, mean=None, rstd=None):
    """ref nn.py:430-437 / unet.py:185-198.  x0 [B,H,W,C0] (+ x1 [B,H,W,C1] concatenated on channels)."""
    _bf16c(x0)
    B = x0.shape[0]
    C0 = x0.shape[-1]
    HW = x0[0].numel() // C0 if B else 0
    C1 = 0
    if x1 is not None:
        _bf16c(x1); C1 = x1.shape[-1]
    Ct = C0 + C1
    out = torch.empty(*x0.shape[:-1], Ct, device=x0.device, dtype=bf16) if out is None else out
    mean = torch.empty(B, 32, device=x0.device, dtype=torch.float32) if mean is None else mean
    rstd = torch.empty(B, 32, device=x0.device, dtype=torch.float32) if rstd is None else rstd
    check(_lib.lib().cdae_gn_fwd(ptr(x0), C0, ptr(x1), C1, B, HW, ptr(gamma), ptr(beta), ptr(film),
                                 film.shape[1] if film is not None else 0, film_off, int(silu), ptr(out), ptr(mean),
                                 ptr(rstd), stream()))
    return out, mean, rstd


def gn_apply_fwd(x0, stats0, gamma, beta, x1=None, stats1=None, film=None, film_off=0, silu=True, out=None, mean=None,
                 rstd=None, ab=None, constants_only=False):
    """gn_fwd as ONE streaming pass: the per-(image, channel) sums stats0 [B,C0,2] (stats1 [B,C1,2]) were accumulated by
    the convolutions that produced x0 (x1) (make_igemm_desc(stats=...)).  constants_only: nothing is streamed - only the
    {a, b} table `ab` [B, C, 2] is written, for a conv that applies the norm while loading (make_igemm_desc(gn=...))."""
    _bf16c(x0); _f32c(stats0)
    B = x0.shape[0]
    C0 = x0.shape[-1]
    HW = x0[0].numel() // C0 if B else 0
    assert tuple(stats0.shape) == (B, C0, 2)
    C1 = 0
    if x1 is not None:
        _bf16c(x1); _f32c(stats1); C1 = x1.shape[-1]
        assert tuple(stats1.shape) == (B, C1, 2)
    Ct = C0 + C1
    if constants_only:
        assert ab is not None and tuple(ab.shape) == (B, Ct, 2)
        out = None
    else:
        out = torch.empty(*x0.shape[:-1], Ct, device=x0.device, dtype=bf16) if out is None else out
    mean = torch.empty(B, 32, device=x0.device, dtype=torch.float32) if mean is None else mean
    rstd = torch.empty(B, 32, device=x0.device, dtype=torch.float32) if rstd is None else rstd
    check(_lib.lib().cdae_gn_apply_fwd(ptr(x0), C0, ptr(stats0), ptr(x1), C1, ptr(stats1), B, HW, ptr(gamma), ptr(beta),
                                       ptr(film), film.shape[1] if film is not None else 0, film_off, int(silu), ptr(out),
                                       ptr(mean), ptr(rstd), ptr(ab), stream()))
    return out, mean, rstd


def gn_bwd(dy, x0, gamma, beta, mean, rstd, x1=None, film=None, film_off=0, silu=True, dx0=None, dx1=None,
           accumulate_dx=0, dgamma=None, dbeta=None, dfilm=None, dadd=None):
    """accumulate_dx: bit0 -> add into dx0, bit1 -> add into dx1 (True == both); dadd: extra bf16 [B,HW,C] gradient.
    The resident cluster kernel: reduces and applies in one launch (layers whose output gradient does not come from a
    statistics-producing data-gradient convolution; otherwise see gn_bwd_apply)."""
    if accumulate_dx is True:
        accumulate_dx = 3
    _bf16c(dy); _bf16c(x0)
    B = x0.shape[0]
    C0 = x0.shape[-1]
    HW = x0[0].numel() // C0 if B else 0
    C1 = 0
    if x1 is not None:
        _bf16c(x1); C1 = x1.shape[-1]
        dx1 = torch.empty_like(x1) if dx1 is None else dx1
    dx0 = torch.empty_like(x0) if dx0 is None else dx0
    check(_lib.lib().cdae_gn_bwd(ptr(dy), ptr(x0), C0, ptr(x1), C1, B, HW, ptr(gamma), ptr(beta), ptr(film),
                                 film.shape[1] if film is not None else 0, film_off, int(silu), ptr(mean), ptr(rstd),
                                 ptr(dadd), ptr(dx0), ptr(dx1), int(accumulate_dx), ptr(dgamma), ptr(dbeta), ptr(dfilm), stream()))
    return dx0, dx1


def gn_bwd_apply(du, x0, gamma, beta, mean, rstd, ws, x1=None, film=None, film_off=0, dx0=None, dx1=None, accumulate_dx=0,
                 dgamma=None, dbeta=None, dfilm=None, dadd=None):
    """GroupNorm backward as ONE streaming pass: du = dy * silu'(u) and ws [B, C, 2] = {sum du, sum du*x} per (sample, channel)
    were produced by the epilogue of the data-gradient conv (make_igemm_desc(gnb=...))."""
    if accumulate_dx is True:
        accumulate_dx = 3
    _bf16c(du); _bf16c(x0); _f32c(ws)
    B = x0.shape[0]
    C0 = x0.shape[-1]
    HW = x0[0].numel() // C0 if B else 0
    C1 = 0
    if x1 is not None:
        _bf16c(x1); C1 = x1.shape[-1]
        dx1 = torch.empty_like(x1) if dx1 is None else dx1
    dx0 = torch.empty_like(x0) if dx0 is None else dx0
    assert ws.numel() == B * 2 * (C0 + C1)
    check(_lib.lib().cdae_gn_bwd_apply(ptr(du), ptr(x0), C0, ptr(x1), C1, B, HW, ptr(gamma), ptr(beta), ptr(film),
                                       film.shape[1] if film is not None else 0, film_off, ptr(mean), ptr(rstd), ptr(ws),
                                       ptr(dadd), ptr(dx0), ptr(dx1), int(accumulate_dx), ptr(dgamma), ptr(dbeta), ptr(dfilm),
                                       stream()))
    return dx0, dx1


# ------------------------------------------------------------------ implicit GEMM (tcgen05)
def conv_segments(chans, ksize=3, transposed=False, wk0=0, src0=0):
    """K segments of a ksize x ksize conv whose input is the channel concat of sources with `chans` channels.
    Weight K index = tap * sum(chans) + channel (OHWI packing).  transposed=True gives the data-gradient taps."""
    ctot = sum(chans)
    segs = []
    taps = [(kh, kw) for kh in range(ksize) for kw in range(ksize)]
    for ti, (kh, kw) in enumerate(taps):
        dh, dw = (kh - ksize // 2, kw - ksize // 2)
        if transposed:
            dh, dw = -dh, -dw
        off = 0
        for si, c in enumerate(chans):
            assert c % 64 == 0, "source channels must be multiples of 64"
            segs.append((src0 + si, dh, dw, 0, c // 64, wk0 + ti * ctot + off))
            off += c
    return segs, wk0 + len(taps) * ctot


def make_igemm_desc(srcs, segs, wgt, out, cout, in_stride=1, bias=None, resid=None, out_mode=0, sps=1, ooh=0, oow=0,
                    bn=0, out_hw=None, bias2=None, stats=None, gnb=None, bias_img=None, gn=None, up2x=False):
    """Fill a cdae_igemm_desc.  srcs: bf16 [N,H,W,C]; wgt: bf16 [rows, K]; out: bf16 NHWC or fp32 NCHW (out_mode 1).
    stats: optional fp32 [N, cout, 2] that the epilogue ACCUMULATES per-(image, channel) sum / sum of squares of the
    stored output into (zero it first) - the GroupNorm statistics of the consumer, see gn_apply_fwd."""
    d = IgemmDesc()
    N, H, W = srcs[0].shape[:3]
    for i, s in enumerate(srcs):
        _bf16c(s)
        assert tuple(s.shape[:3]) == (N, H, W)
        d.src[i] = s.data_ptr()
        d.src_c[i] = s.shape[3]
    d.nsrc, d.N, d.H, d.W, d.in_stride = len(srcs), N, H, W, in_stride
    assert len(segs) <= _lib.MAX_SEG
    d.nseg = len(segs)
    for i, sg in enumerate(segs):
        d.seg[i] = Seg(*sg)
    _bf16c(wgt)
    d.wgt, d.wrows, d.wk = wgt.data_ptr(), wgt.shape[0], wgt.shape[1]
    d.out, d.out_mode = out.data_ptr(), out_mode
    if out_mode == 0:
        OH, OW = (out.shape[1], out.shape[2]) if out_hw is None else out_hw
        d.ldo = out.shape[-1]
    elif out_mode == 2:         # fp32 [N, OH, OW, ldo] row-major
        assert out.dtype == torch.float32 and out.is_contiguous()
        OH, OW = out.shape[1], out.shape[2]
        d.ldo = out.shape[-1]
    else:
        OH, OW = out.shape[2], out.shape[3]
        d.ldo = 0
    d.OH, d.OW, d.cout = OH, OW, cout
    d.sps, d.ooh, d.oow = sps, ooh, oow
    d.bias = ptr(bias)
    d.bias2 = ptr(bias2)
    d.resid = ptr(resid)
    d.ldr = resid.shape[-1] if resid is not None else 0
    d.bn = bn
    if stats is not None:
        _f32c(stats)
        assert out_mode == 0 and tuple(stats.shape) == (N, cout, 2)
    d.stats = ptr(stats)
    if gnb is not None:
        # GroupNorm-backward fusion of a data-gradient launch: gnb = dict(x0=, x1=None, ab=None, ws=, silu=) - the epilogue
        # stores du = dy * silu'(u) and accumulates ws [N, cout, 2] += {sum du, sum du*x} (see include/cdae.h)
        x0, x1, ws, ab = gnb["x0"], gnb.get("x1"), gnb["ws"], gnb.get("ab")
        _bf16c(x0); _f32c(ws)
        assert out_mode == 0 and tuple(ws.shape) == (N, cout, 2) and x0.shape[3] + (x1.shape[3] if x1 is not None else 0) == cout
        d.gnb_ws, d.gnb_ab, d.gnb_x0, d.gnb_x1 = ptr(ws), ptr(ab), ptr(x0), ptr(x1)
        d.gnb_c0, d.gnb_ld0, d.gnb_ld1 = x0.shape[3], x0.shape[3], (x1.shape[3] if x1 is not None else 0)
        d.gnb_silu = int(bool(gnb.get("silu", True)))
    if bias_img is not None:       # fp32 [N, >= cout] view (row pitch = stride(0)): + bias_img[n, co] in the epilogue
        assert bias_img.dtype == torch.float32 and bias_img.stride(1) == 1 and bias_img.shape[0] == N
        d.bias_img, d.bias_img_ld = bias_img.data_ptr(), bias_img.stride(0)
    d.up2x = int(bool(up2x))          # sources at half the output resolution, nearest x2 upsampling applied on load
    for i in range(4):
        d.gn_off[i] = -1
    if gn is not None:
        # GroupNorm(+FiLM)+SiLU applied to sources on load: gn = (ab [N, C, 2] from gn_apply_fwd(constants_only=True),
        # [table column of source i's channel 0, or -1 for a source that is taken raw])
        gab, offs = gn
        _f32c(gab)
        assert gab.dim() == 3 and gab.shape[0] == N and gab.shape[2] == 2 and len(offs) == len(srcs)
        d.gn_ab, d.gn_c = gab.data_ptr(), gab.shape[1]
        for i, o in enumerate(offs):
            d.gn_off[i] = int(o)
    d._keep = (srcs, wgt, out, bias, bias2, resid, stats, gnb, bias_img, gn)   # keep tensors alive as long as the descriptor
    return d


PROFILE = None   # set to a list to record (kind, info, flops, start_event, end_event) per tensor-core launch (tools/)


def _profiled(kind, info, flops, launch):
    if PROFILE is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    PROFILE.append((kind, info, flops, e0, e1))


def igemm(desc):
    if PROFILE is not None:
        s = max(desc.in_stride, 1)
        pix = desc.N * ((desc.H + s - 1) // s) * ((desc.W + s - 1) // s)
        k = 64 * sum(desc.seg[i].nchunk for i in range(desc.nseg))
        info = f"N{desc.N} {desc.H}x{desc.W} s{s} K{k} cout{desc.cout} nsrc{desc.nsrc}"
        return _profiled("igemm", info, 2.0 * pix * k * desc.cout,
                         lambda: check(_lib.lib().cdae_igemm(C.byref(desc), stream())))
    check(_lib.lib().cdae_igemm(C.byref(desc), stream()))


def make_wgrad_desc(dy, src, dw, cout, cin, ksize=3, in_stride=1, c0=0, ci_off=0, cin_real=None, dw_ld=None, splits=0,
                    dbias=None):
    """dw (fp32 [cout, taps, dw_ld]) += dy^T * shifted(src).  dy: bf16 [N,OH,OW,ldy]; src: bf16 [N,H,W,C]."""
    _bf16c(dy); _bf16c(src)
    d = WgradDesc()
    d.dy, d.ldy, d.cout = dy.data_ptr(), dy.shape[-1], cout
    d.src, d.src_c, d.c0, d.cin = src.data_ptr(), src.shape[-1], c0, cin
    d.N, d.H, d.W = src.shape[0], src.shape[1], src.shape[2]
    d.OH, d.OW = dy.shape[1], dy.shape[2]
    d.in_stride, d.ksize = in_stride, ksize
    d.dw = dw.data_ptr()
    d.dw_ld = dw_ld if dw_ld is not None else cin
    d.ci_off, d.cin_real, d.splits = ci_off, (cin_real if cin_real is not None else cin), splits
    d.dbias = ptr(dbias)
    d._keep = (dy, src, dw, dbias)
    return d


def wgrad(desc):
    if PROFILE is not None:
        pix = desc.N * desc.OH * desc.OW
        info = f"N{desc.N} {desc.OH}x{desc.OW} s{desc.in_stride} k{desc.ksize} cin{desc.cin} cout{desc.cout}"
        return _profiled("wgrad", info, 2.0 * pix * desc.cin * desc.cout * desc.ksize ** 2,
                         lambda: check(_lib.lib().cdae_wgrad(C.byref(desc), stream())))
    check(_lib.lib().cdae_wgrad(C.byref(desc), stream()))


# ------------------------------------------------------------------ attention
def attn_fwd(qkv, heads, out=None, lse=None):
    """ref unet.py:239-253. qkv bf16 [B, T, 3C] with per-head [q|k|v] interleave; returns out bf16 [B,T,C], lse fp32 [B,heads,T]."""
    _bf16c(qkv)
    B, T, C3 = qkv.shape
    Cc = C3 // 3
    ch = Cc // heads
    out = torch.empty(B, T, Cc, device=qkv.device, dtype=bf16) if out is None else out
    lse = torch.empty(B, heads, T, device=qkv.device, dtype=torch.float32) if lse is None else lse
    check(_lib.lib().cdae_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, T, heads, ch, stream()))
    return out, lse


def attn_bwd(qkv, out, dout, lse, heads, dqkv=None, dsum=None):
    """dsum: fp32 [B, heads, T] scratch (rowsum(dout*out)); pass a preallocated buffer on graph-captured paths."""
    _bf16c(qkv); _bf16c(out); _bf16c(dout)
    B, T, C3 = qkv.shape
    ch = C3 // 3 // heads
    dqkv = torch.empty_like(qkv) if dqkv is None else dqkv
    dsum = torch.empty(B, heads, T, device=qkv.device, dtype=torch.float32) if dsum is None else _f32c(dsum)
    check(_lib.lib().cdae_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dsum), ptr(dqkv), B, T, heads, ch, stream()))
    return dqkv


# ------------------------------------------------------------------ causal DAG mask layer
def dag_fwd(u, A, param_ptrs, n, d, D, out=None):
    """ref nn.py:290-312 in one launch. u fp32 [B, n*d]; A fp32 [n, n]; param_ptrs: device int64 [4n] -> z_post [B, n*d]"""
    _f32c(u); _f32c(A)
    B = u.shape[0]
    z = torch.empty_like(u) if out is None else out
    check(_lib.lib().cdae_dag_fwd(ptr(u), ptr(A), ptr(param_ptrs), ptr(z), B, n, d, D, stream()))
    return z


def dag_bwd(u, A, param_ptrs, dz, grad_ptrs, ws, n, d, D, du=None, du_add=None):
    """-> du (+ du_add); parameter gradients are accumulated through grad_ptrs (device int64 [4n]); ws: zeroed fp32 [B, n*d]"""
    _f32c(u); _f32c(A); _f32c(dz); _f32c(ws)
    B = u.shape[0]
    du = torch.empty_like(u) if du is None else du
    check(_lib.lib().cdae_dag_bwd(ptr(u), ptr(A), ptr(param_ptrs), ptr(dz), ptr(grad_ptrs), ptr(ws), ptr(du), ptr(du_add), B,
                                  n, d, D, stream()))
    return du


# ------------------------------------------------------------------ fp32 representation path (csrc/rep.cu)
def sgemm(C_, A, B_, M, N, K, a_st, b_st, c_st, a_mode=0, b_mode=0, c_mode=0, bias=None, act_out=0, colstats=None, splits=0,
          geo=None):
    """C[M,N] (=, +=, scatter+=) A[M,K] B[K,N]; *_st = element strides (row, col) of each operand view; geo = (sb, sc, sh, sw,
    Cin, H, W, OH, OW, ab) for the im2col modes (see include/cdae.h)."""
    d = SgemmDesc()
    d.A, d.a_sm, d.a_sk, d.a_mode = A.data_ptr(), a_st[0], a_st[1], a_mode
    d.B, d.b_sk, d.b_sn, d.b_mode = B_.data_ptr(), b_st[0], b_st[1], b_mode
    d.C, d.c_sm, d.c_sn, d.c_mode = C_.data_ptr(), c_st[0], c_st[1], c_mode
    d.bias, d.act_out, d.colstats = ptr(bias), act_out, ptr(colstats)
    d.M, d.N, d.K, d.splits = M, N, K, splits
    if geo is not None:
        d.g_sb, d.g_sc, d.g_sh, d.g_sw, d.g_cin, d.g_h, d.g_w, d.g_oh, d.g_ow = geo[:9]
        d.g_ab = ptr(geo[9])
    check(_lib.lib().cdae_sgemm(C.byref(d), stream()))


_ONE = {}


def _one(device):
    t = _ONE.get(str(device))
    if t is None:
        t = _ONE[str(device)] = torch.ones(4, device=device)
    return t


def linear_fwd(x, W, b, out, silu_in=False, accumulate=False, act_out=0):
    """out (=|+=) f(x) W^T + b  (nn.Linear; f = SiLU when silu_in).  x [B,K], W [N,K], out [B,N] contiguous fp32."""
    Bn, K = x.shape
    N = W.shape[0]
    sgemm(out, x, W, Bn, N, K, (K, 1), (1, K), (N, 1), a_mode=1 if silu_in else 0, c_mode=1 if accumulate else 0, bias=b,
          act_out=act_out, splits=1)
    return out


def linear_bwd(x, W, dy, dW, db, dx=None, silu_in=False, dx_accumulate=False):
    """gradients of y = f(x) W^T + b: dW += dy^T f(x), db += colsum(dy) (straight into the flat gradient arena views),
    dx (=|+=) dy W - the caller applies f' (silu_bwd) when silu_in."""
    Bn, K = x.shape
    N = W.shape[0]
    if dW is not None:
        sgemm(dW, dy, x, N, K, Bn, (1, N), (K, 1), (K, 1), b_mode=1 if silu_in else 0, c_mode=1)
    if db is not None:
        sgemm(db, _one(x.device), dy, 1, N, Bn, (0, 0), (N, 1), (N, 1), c_mode=1)
    if dx is not None:
        if not dx_accumulate:
            zero_(dx)
        sgemm(dx, dy, W, Bn, K, N, (N, 1), (K, 1), (K, 1), c_mode=1)
    return dx


def timestep_embedding(t, freqs, out, tmap=None, scale=0.0):
    """t: [B] int64 / fp32, or ONE element (a device-side loop counter) broadcast over the B rows of `out`"""
    B, dim = out.shape
    assert t.is_cuda and t.dtype in (torch.int64, torch.float32) and t.is_contiguous() and t.numel() in (1, B)
    check(_lib.lib().cdae_timestep_embedding(ptr(t), int(t.dtype == torch.float32), 1 if t.numel() == B else 0,
                                             ptr(tmap), float(scale), ptr(freqs), ptr(out), B, dim, stream()))
    return out


def step_tick(step64, step32, delta):
    check(_lib.lib().cdae_step_tick(ptr(step64), ptr(step32), int(delta), stream()))


def randn_(out, state, bernoulli=False, keep_prob=0.5):
    """fill `out` (fp32) with N(0,1) draws, or Bernoulli(keep_prob) 0/1; state: device int64[2] {seed, offset}"""
    check(_lib.lib().cdae_randn(ptr(_f32c(out)), out.numel(), ptr(state), int(bernoulli), float(keep_prob), stream()))
    return out


def silu_cast(x, out, silu=True):
    """out (bf16) = SiLU(x) (or x)"""
    check(_lib.lib().cdae_silu_cast(ptr(_f32c(x)), ptr(_bf16c(out)), x.numel(), int(silu), stream()))
    return out


def silu_bwd_(g, x):
    check(_lib.lib().cdae_silu_bwd(ptr(_f32c(g)), ptr(_f32c(x)), g.numel(), stream()))
    return g


def softplus_bwd_(g, var):
    check(_lib.lib().cdae_softplus_bwd(ptr(_f32c(g)), ptr(_f32c(var)), g.numel(), stream()))
    return g


def embed_rows_(x, table, idx, backward=False, dtable=None):
    check(_lib.lib().cdae_embed_rows(ptr(_f32c(x)), ptr(table), ptr(idx), x.shape[0], x.shape[1], int(backward), ptr(dtable),
                                     stream()))
    return x


def bn_finalize(stats, count, bn, train, ab, ms):
    check(_lib.lib().cdae_bn_finalize(ptr(stats), float(count), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
                                      ptr(bn.running_var), ptr(bn.num_batches_tracked), int(train), ptr(ab), ptr(ms),
                                      bn.weight.shape[0], stream()))


def enc_head(src, ab, out, B, P, Cc, backward=False):
    check(_lib.lib().cdae_enc_head(ptr(src), ptr(ab), ptr(out), B, P, Cc, int(backward), stream()))
    return out


def bn_lrelu_bwd_(dact, raw, ab, ms, gamma, sums, dgamma, dbeta, M, Cc):
    check(_lib.lib().cdae_bn_lrelu_bwd(ptr(dact), ptr(raw), ptr(ab), ptr(ms), ptr(gamma), ptr(sums), ptr(dgamma), ptr(dbeta),
                                       M, Cc, stream()))
    return dact


def latent_fwd(mu, var, zp, xi, keep, c, z, zp_out, kld, n, causal, var_scale):
    B, D = mu.shape
    check(_lib.lib().cdae_latent_fwd(ptr(mu), ptr(var), ptr(zp), ptr(xi), ptr(keep), ptr(c), ptr(z), ptr(zp_out), ptr(kld),
                                     B, D, n, int(causal), float(var_scale), stream()))


def latent_bwd(mu, var, zp, xi, keep, c, dz, dkld, dzp_ext, dmu_ext, dvar_ext, dzp, dmu, dvar, n, causal, var_scale):
    B, D = mu.shape
    check(_lib.lib().cdae_latent_bwd(ptr(mu), ptr(var), ptr(zp), ptr(xi), ptr(keep), ptr(c), ptr(dz), ptr(dkld), ptr(dzp_ext),
                                     ptr(dmu_ext), ptr(dvar_ext), ptr(dzp), ptr(dmu), ptr(dvar), B, D, n, int(causal),
                                     float(var_scale), stream()))


def step_loss(mse, kld, keep, w, kl_weight, t, num_timesteps, loss, gscale, dkld, total, logsums):
    check(_lib.lib().cdae_step_loss(ptr(mse), ptr(kld), ptr(keep), ptr(w), ptr(kl_weight), ptr(t), num_timesteps,
                                    mse.shape[0], ptr(loss), ptr(gscale), ptr(dkld), ptr(total), ptr(logsums), stream()))

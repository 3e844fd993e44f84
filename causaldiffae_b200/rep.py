"""The [B, 512]-sized representation path as launch sequences over csrc/rep.cu (no ATen / cuBLAS / cuDNN kernels):

  TrunkRunner    timestep_embedding -> time_embed (+ label_emb, c_emb) + up_emb(z) -> all FiLM projections
                 (ref nn.py:551-569, unet.py:545-554,616,148-154)
  EncoderRunner  GaussianConvEncoder.encode: [conv3x3 s2 -> BatchNorm2d -> LeakyReLU] x L -> fc_mu / softplus(fc_var)+1e-8
                 (ref nn.py:93-110), training (batch statistics, running buffers advanced) and inference
  latent         reparameterize + classifier-free keep mask (+ the closed-form KL of representation_loss)
                 (ref nn.py:460-467, unet.py:590-613, gaussian_diffusion.py:718-766)

Each runner has an explicit forward / backward over a state object of preallocated buffers: the fused training step
(train_util.FusedStep) replays them inside one CUDA graph, the autograd Functions below wrap the same launches for callers
of the module API (`model(x, t, ...)`, `model.rep_emb.encode`).  Parameter gradients never pass through autograd: the
backward launches accumulate straight into `param.grad` (views of the engine's flat gradient arena)."""
import math
import os

import torch as th

from . import ops

# weight gradients of the conv encoder and the bf16 weight pack of the fused step run on the side stream (ops.on_side)
REP_SIDE_STREAM = os.environ.get("CDAE_REP_SIDE_STREAM", "1") != "0"

_FREQS = {}


def freqs_for(dim, device, max_period=10000):
    """ref nn.py:562-564, computed on the host exactly like the reference and uploaded once"""
    half = dim // 2
    key = (half, max_period, str(device))
    f = _FREQS.get(key)
    if f is None:
        f = th.exp(-math.log(max_period) * th.arange(start=0, end=half, dtype=th.float32) / half).to(device)
        _FREQS[key] = f
    return f


_ANCHORS = {}


def anchor(device):
    """A leaf that requires grad: custom Functions whose tensor inputs do not (timesteps, images) still have to be recorded
    by autograd, because their backward is what accumulates the parameter gradients."""
    a = _ANCHORS.get(str(device))
    if a is None:
        a = _ANCHORS[str(device)] = th.zeros((), device=device, requires_grad=True)
    return a


def _g(p):
    """gradient buffer of a parameter (a view of the flat gradient arena once the engine exists)"""
    if p.grad is None:
        p.grad = th.zeros_like(p)
    return p.grad


class _State:
    pass


# ---------------------------------------------------------------------------------------------------- embedding trunk
class TrunkRunner:
    def __init__(self, model):
        self.m = model

    def alloc(self, B, device, train):
        m = self.m
        st = _State()
        ted = m.model_channels * 4
        f = lambda *s: th.empty(*s, device=device, dtype=th.float32)   # noqa: E731
        st.temb, st.h1, st.emb = f(B, m.model_channels), f(B, ted), f(B, ted)
        st.a_bf = th.empty(B, ted, device=device, dtype=th.bfloat16)          # SiLU(emb): A operand of the FiLM GEMM
        st.hc = f(B, 256) if m.c_dim is not None else None
        if train:
            st.demb, st.dh1 = f(B, ted), f(B, ted)
            st.dfilm_bf = th.empty(B, m.engine.film_width, device=device, dtype=th.bfloat16)
            st.dhc = f(B, 256) if m.c_dim is not None else None
        return st

    def forward(self, st, t, y, c, z, film_out, tmap=None, scale=0.0):
        m, eng = self.m, self.m.engine
        ops.timestep_embedding(t, freqs_for(m.model_channels, t.device), st.temb, tmap, scale)
        te0, te2 = m.time_embed[0], m.time_embed[2]
        ops.linear_fwd(st.temb, te0.weight, te0.bias, st.h1)
        ops.linear_fwd(st.h1, te2.weight, te2.bias, st.emb, silu_in=True)
        if m.num_classes is not None:
            ops.embed_rows_(st.emb, m.label_emb.weight, y)
        if m.c_dim is not None:
            ops.linear_fwd(c, m.c_emb[0].weight, m.c_emb[0].bias, st.hc)
            ops.linear_fwd(st.hc, m.c_emb[2].weight, m.c_emb[2].bias, st.emb, silu_in=True, accumulate=True)
        if z is not None:
            ops.linear_fwd(z, m.up_emb.weight, m.up_emb.bias, st.emb, accumulate=True)
        eng.film_fwd(st.emb, film_out, st.a_bf)
        return film_out

    def backward(self, st, y, c, z, dfilm, dz_out=None, part="all"):
        """part: "film" = only the FiLM projection's backward (after it every gradient in [0, engine.early_end) is final),
        "rest" = what follows it, "all" = both"""
        m, eng = self.m, self.m.engine
        if part != "rest":
            eng.film_bwd(st.emb, st.a_bf, dfilm, st.dfilm_bf, st.demb)
        if part == "film":
            return
        if z is not None:
            ops.linear_bwd(z, m.up_emb.weight, st.demb, _g(m.up_emb.weight), _g(m.up_emb.bias), dx=dz_out)
        if m.c_dim is not None:
            ops.linear_bwd(st.hc, m.c_emb[2].weight, st.demb, _g(m.c_emb[2].weight), _g(m.c_emb[2].bias), dx=st.dhc, silu_in=True)
            ops.silu_bwd_(st.dhc, st.hc)
            ops.linear_bwd(c, m.c_emb[0].weight, st.dhc, _g(m.c_emb[0].weight), _g(m.c_emb[0].bias))
        if m.num_classes is not None:
            ops.embed_rows_(st.demb, None, y, backward=True, dtable=_g(m.label_emb.weight))
        te0, te2 = m.time_embed[0], m.time_embed[2]
        ops.linear_bwd(st.h1, te2.weight, st.demb, _g(te2.weight), _g(te2.bias), dx=st.dh1, silu_in=True)
        ops.silu_bwd_(st.dh1, st.h1)
        ops.linear_bwd(st.temb, te0.weight, st.dh1, _g(te0.weight), _g(te0.bias))


class _FilmFn(th.autograd.Function):
    """film = emb_layers(SiLU(time_embed(temb(t)) [+ label_emb(y)] [+ c_emb(c)] [+ up_emb(z)])) for all ResBlocks at once"""

    @staticmethod
    def forward(ctx, model, _anchor, t, y, c, z, tmap, scale, train):
        run = model.engine.trunk                   # train = th.is_grad_enabled() at the call site (it is off in here)
        model.engine.pack()                        # the FiLM projection reads the packed bf16 weights
        B = t.shape[0]
        if t.dtype not in (th.int64, th.float32):
            t = t.float() if t.is_floating_point() else t.long()
        st = run.alloc(B, t.device, train)
        zc = None if z is None else z.detach().float().contiguous()
        cc = None if (c is None or model.c_dim is None) else c.detach().float().contiguous()
        film = th.empty(B, model.engine.film_width, device=t.device, dtype=th.float32)
        run.forward(st, t.contiguous(), y, cc, zc, film, tmap, scale)
        ctx.model, ctx.st, ctx.y, ctx.c, ctx.z = model, st, y, cc, zc
        ctx.z_needs = z is not None and z.requires_grad
        return film

    @staticmethod
    def backward(ctx, dfilm):
        m = ctx.model
        dz = th.empty_like(ctx.z) if ctx.z is not None else None
        m.engine.trunk.backward(ctx.st, ctx.y, ctx.c, ctx.z, dfilm.float().contiguous(), dz)
        return None, None, None, None, None, (dz if ctx.z_needs else None), None, None, None


# ---------------------------------------------------------------------------------------------------- conv encoder
class EncoderRunner:
    def __init__(self, enc):
        self.enc = enc

    def geometry(self, Cin, H, W):
        geo = []
        for blk in self.enc.encoder:
            co = blk[0].out_channels
            oh, ow = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            geo.append((Cin, co, H, W, oh, ow))
            Cin, H, W = co, oh, ow
        return geo

    def alloc(self, B, Cin, H, W, device, train):
        st = _State()
        st.geo = self.geometry(Cin, H, W)
        f = lambda *s: th.empty(*s, device=device, dtype=th.float32)   # noqa: E731
        st.raw = [f(B, oh, ow, co) for (_, co, _, _, oh, ow) in st.geo]
        st.stats = [th.zeros(co, 2, device=device, dtype=th.float64) for (_, co, _, _, _, _) in st.geo]
        st.ab = [f(co, 2) for (_, co, _, _, _, _) in st.geo]
        st.ms = [f(co, 2) for (_, co, _, _, _, _) in st.geo]
        _, co, _, _, oh, ow = st.geo[-1]
        st.P, st.C = oh * ow, co
        D = self.enc.latent_dim
        st.hfeat, st.mu, st.var = f(B, co * oh * ow), f(B, D), f(B, D)
        if train:
            st.dact = [f(B, oh_, ow_, co_) for (_, co_, _, _, oh_, ow_) in st.geo]
            st.dh = f(B, co * oh * ow)
            st.sums = th.zeros(max(g[1] for g in st.geo), 2, device=device, dtype=th.float64)
            st.dpre = f(B, D)
        return st

    @staticmethod
    def _src(st, x, l):
        """(tensor, (sb, sc, sh, sw), ab) of the input of layer l: the NCHW image, or the NHWC raw output of layer l-1"""
        ci, _, h, w, _, _ = st.geo[l]
        if l == 0:
            return x, (ci * h * w, h * w, w, 1), None
        return st.raw[l - 1], (h * w * ci, 1, w * ci, ci), st.ab[l - 1]

    def features(self, st, x, train):
        """the conv stack: [conv3x3 s2 -> BatchNorm2d -> LeakyReLU] x L -> th.flatten  (ref nn.py:100-104) -> st.hfeat [B, C*P]"""
        B = x.shape[0]
        for l, (ci, co, h, w, oh, ow) in enumerate(st.geo):
            conv, bn = self.enc.encoder[l][0], self.enc.encoder[l][1]
            src, strides, ab = self._src(st, x, l)
            M, K = B * oh * ow, ci * 9
            if train:
                ops.zero_(st.stats[l])
            ops.sgemm(st.raw[l], src, conv.weight, M, co, K, (0, 0), (1, K), (co, 1), a_mode=2, bias=conv.bias,
                      colstats=st.stats[l] if train else None, splits=1, geo=(*strides, ci, h, w, oh, ow, ab))
            ops.bn_finalize(st.stats[l], M, bn, train, st.ab[l], st.ms[l])
        ops.enc_head(st.raw[-1], st.ab[-1], st.hfeat, B, st.P, st.C)
        return st.hfeat

    def forward(self, st, x, train):
        self.features(st, x, train)
        e = self.enc
        ops.linear_fwd(st.hfeat, e.fc_mu.weight, e.fc_mu.bias, st.mu)
        ops.linear_fwd(st.hfeat, e.fc_var.weight, e.fc_var.bias, st.var, act_out=1)
        return st.mu, st.var

    def backward(self, st, x, dmu, dvar):
        """dmu / dvar: gradients w.r.t. mu and var = softplus(.) + 1e-8 (dvar is consumed: overwritten in place)"""
        e = self.enc
        ops.softplus_bwd_(dvar, st.var)
        ops.linear_bwd(st.hfeat, e.fc_mu.weight, dmu, _g(e.fc_mu.weight), _g(e.fc_mu.bias), dx=st.dh)
        ops.linear_bwd(st.hfeat, e.fc_var.weight, dvar, _g(e.fc_var.weight), _g(e.fc_var.bias), dx=st.dh, dx_accumulate=True)
        self.features_backward(st, x)

    def features_backward(self, st, x):
        """st.dh (gradient w.r.t. the flattened features) -> every conv / BatchNorm parameter gradient of the stack"""
        e, B = self.enc, x.shape[0]
        ops.enc_head(st.dh, None, st.dact[-1], B, st.P, st.C, backward=True)
        for l in reversed(range(len(st.geo))):
            ci, co, h, w, oh, ow = st.geo[l]
            conv, bn = e.encoder[l][0], e.encoder[l][1]
            M, K = B * oh * ow, ci * 9
            ops.bn_lrelu_bwd_(st.dact[l], st.raw[l], st.ab[l], st.ms[l], bn.weight, st.sums, _g(bn.weight), _g(bn.bias), M, co)
            src, strides, ab = self._src(st, x, l)
            # dW[co][k] += sum_pixels draw[m][co] * col[m][k]  (rows = conv k through the transposed im2col view); nothing
            # on the chain reads it: it runs beside the data gradient of the same layer (both are small-grid launches)
            gw = _g(conv.weight)
            if REP_SIDE_STREAM:
                with ops.on_side():
                    ops.sgemm(gw, src, st.dact[l], K, co, M, (0, 0), (co, 1), (1, K), a_mode=3, c_mode=1,
                              geo=(*strides, ci, h, w, oh, ow, ab))
            else:
                ops.sgemm(gw, src, st.dact[l], K, co, M, (0, 0), (co, 1), (1, K), a_mode=3, c_mode=1,
                          geo=(*strides, ci, h, w, oh, ow, ab))
            _g(conv.bias)   # feeds a batch-statistics BatchNorm: its gradient is identically zero (sum of draw per channel)
            if l > 0:
                ops.zero_(st.dact[l - 1])
                ops.sgemm(st.dact[l - 1], st.dact[l], conv.weight, M, K, co, (co, 1), (K, 1), (0, 0), c_mode=2,
                          geo=(h * w * ci, 1, w * ci, ci, ci, h, w, oh, ow, None))
        if REP_SIDE_STREAM:
            ops.join_side()


class _EncodeFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, enc, _anchor, x, want_bwd):
        run = enc.runner
        xc = x.detach().float().contiguous()
        train_stats = enc.training
        if want_bwd and not train_stats:
            raise NotImplementedError("gradients through the conv encoder in eval mode (running BatchNorm statistics) are not "
                                      "built: call it under no_grad, or in train mode")
        st = run.alloc(xc.shape[0], xc.shape[1], xc.shape[2], xc.shape[3], xc.device, want_bwd)
        mu, var = run.forward(st, xc, train_stats)
        ctx.enc, ctx.st, ctx.x = enc, st, xc
        return mu, var

    @staticmethod
    def backward(ctx, dmu, dvar):
        st = ctx.st
        dmu = th.zeros_like(st.mu) if dmu is None else dmu.float().contiguous()
        dv = th.zeros_like(st.var) if dvar is None else dvar.float().clone().contiguous()
        ctx.enc.runner.backward(st, ctx.x, dmu, dv)
        return None, None, None, None


# ---------------------------------------------------------------------------------------------------- latent
class _LatentFn(th.autograd.Function):
    """z = (zp + sqrt(var_scale var) xi) keep, zp_masked = zp keep   (ref nn.py:460-467, unet.py:590-613)"""

    @staticmethod
    def forward(ctx, var, zp, xi, keep, var_scale, n):
        var_c, zp_c, xi_c = var.detach().float().contiguous(), zp.detach().float().contiguous(), xi.float().contiguous()
        keep_c = None if keep is None else keep.float().contiguous()
        z, zpm = th.empty_like(zp_c), th.empty_like(zp_c)
        ops.latent_fwd(zp_c, var_c, zp_c, xi_c, keep_c, None, z, zpm, None, n, False, var_scale)
        ctx.save_for_backward(var_c, zp_c, xi_c, keep_c if keep_c is not None else th.empty(0, device=zp_c.device))
        ctx.has_keep, ctx.var_scale, ctx.n = keep_c is not None, var_scale, n
        return z, zpm

    @staticmethod
    def backward(ctx, dz, dzpm):
        var, zp, xi, keep = ctx.saved_tensors
        keep = keep if ctx.has_keep else None
        dzp, dmu, dvar = th.empty_like(zp), th.empty_like(zp), th.empty_like(zp)
        ops.latent_bwd(zp, var, zp, xi, keep, None, None if dz is None else dz.float().contiguous(), None,
                       None if dzpm is None else dzpm.float().contiguous(), None, None, dzp, dmu, dvar, ctx.n, False,
                       ctx.var_scale)
        return dvar, dzp, None, None, None, None

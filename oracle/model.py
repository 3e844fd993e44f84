"""Oracle: functional fp32 restatement of the CausalDiffAE UNet (reference improved_diffusion/unet.py, nn.py).

Test infrastructure only (see oracle/__init__.py).  The network is evaluated directly from a
reference-format ``state_dict`` (key names of SURVEY.md Appendix F) with plain torch ops in fp32;
autograd supplies the backward.  Layer helpers are exposed individually so parity tests can be
teacher-forced per layer.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import schedules


@dataclass
class UNetConfig:
    image_size: int = 64
    in_channels: int = 3
    model_channels: int = 128
    out_channels: int = 3
    num_res_blocks: int = 2
    attention_ds: Tuple[int, ...] = (4, 8)
    channel_mult: Tuple[int, ...] = (1, 2, 3, 4)
    num_heads: int = 4
    num_heads_upsample: int = -1
    use_scale_shift_norm: bool = True
    num_classes: Optional[int] = None
    c_dim: Optional[int] = None
    rep_dim: Optional[int] = None
    n_vars: int = 4
    causal_modeling: bool = False
    masking: bool = False
    drop_prob: float = 0.5
    A: Optional[list] = None            # oracle patch 2: injectable DAG (default = ref unet.py:571-575)
    encoder_dims: list = field(default_factory=list)

    def __post_init__(self):
        if self.num_heads_upsample == -1:
            self.num_heads_upsample = self.num_heads
        if self.rep_dim is not None and not self.encoder_dims:
            self.encoder_dims = schedules.encoder_hidden_dims(self.image_size, self.n_vars)
        if self.A is None:
            self.A = schedules.default_dag(self.n_vars)


def config_from_flags(image_size=64, num_channels=128, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
                      attention_resolutions="16,8", learn_sigma=False, class_cond=False, use_scale_shift_norm=True,
                      context_cond=False, rep_cond=False, n_vars=4, causal_modeling=False, in_channels=3,
                      masking=False, rep_dim=512, A=None, **_):
    """create_model (ref script_util.py:119-179)."""
    return UNetConfig(
        image_size=image_size, in_channels=in_channels, model_channels=num_channels,
        out_channels=in_channels * (2 if learn_sigma else 1), num_res_blocks=num_res_blocks,
        attention_ds=schedules.attention_ds_for(image_size, attention_resolutions),
        channel_mult=schedules.channel_mult_for(image_size), num_heads=num_heads,
        num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
        num_classes=10 if class_cond else None, c_dim=4 if context_cond else None,
        rep_dim=rep_dim if rep_cond else None, n_vars=n_vars, causal_modeling=causal_modeling,
        masking=masking, A=A)


# ----------------------------------------------------------------------------- primitives
def silu(x):
    """ref nn.py:430-432."""
    return x * torch.sigmoid(x)


def group_norm32(x, w, b):
    """GroupNorm32 (ref nn.py:435-437, 541-548): 32 groups, eps 1e-5, fp32 statistics."""
    return F.group_norm(x.float(), 32, w, b, eps=1e-5).type(x.dtype)


def timestep_embedding(t, dim, max_period=10000):
    """ref nn.py:551-569: [cos(t f_k), sin(t f_k)], f_k = exp(-ln(max_period) k / half), cos half first."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def resblock(sd, p, x, emb, use_scale_shift_norm=True, drop=None):
    """ResBlock._forward (ref unet.py:185-198). `p` is the key prefix, e.g. 'input_blocks.1.0.'.  `drop`: the multiplier
    nn.Dropout applies between SiLU and the second conv (mask / (1 - p)), injected so that a test can use the masks of the
    implementation under test (torch's own dropout stream cannot be reproduced by another generator)."""
    h = group_norm32(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"])
    h = F.conv2d(silu(h), sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"]).type(h.dtype)[..., None, None]
    if use_scale_shift_norm:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = group_norm32(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"]) * (1 + scale) + shift
    else:
        h = group_norm32(h + e, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"])
    h = silu(h)
    if drop is not None:
        h = h * drop
    h = F.conv2d(h, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        w = sd[p + "skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + "skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def qkv_attention(qkv):
    """QKVAttention.forward (ref unet.py:239-253): qkv [N, 3*ch, T] -> [N, ch, T]; scale ch^-1/4 on q and k."""
    ch = qkv.shape[1] // 3
    q, k, v = torch.split(qkv, ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * s, k * s)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    return torch.einsum("bts,bcs->bct", w, v)


def attention_block(sd, p, x, num_heads):
    """AttentionBlock._forward (ref unet.py:223-231)."""
    b, c, *sp = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(group_norm32(xf, sd[p + "norm.weight"], sd[p + "norm.bias"]), sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    h = qkv_attention(qkv.reshape(b * num_heads, -1, qkv.shape[2])).reshape(b, -1, qkv.shape[2])
    h = F.conv1d(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return (xf + h).reshape(b, c, *sp)


def upsample(sd, p, x):
    """Upsample.forward (ref unet.py:69-79): nearest x2 then 3x3 conv."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)


def downsample(sd, p, x):
    """Downsample.forward (ref unet.py:103-105): 3x3 stride-2 conv."""
    return F.conv2d(x, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)


def encoder_encode(sd, cfg, x, training=True, p="rep_emb."):
    """GaussianConvEncoder.encode (ref nn.py:93-110): [conv3x3 s2 -> BatchNorm2d -> LeakyReLU(0.01)] x L ->
    flatten -> fc_mu, softplus(fc_var) + 1e-8.  BatchNorm uses batch statistics in training mode (Q15);
    running-stat updates are a side effect on `sd` buffers exactly as nn.BatchNorm2d does (momentum 0.1)."""
    h = x
    for k in range(len(cfg.encoder_dims)):
        q = f"{p}encoder.{k}."
        h = F.conv2d(h, sd[q + "0.weight"], sd[q + "0.bias"], stride=2, padding=1)
        rm, rv = sd.get(q + "1.running_mean"), sd.get(q + "1.running_var")
        h = F.batch_norm(h, rm, rv, sd[q + "1.weight"], sd[q + "1.bias"], training=training, momentum=0.1, eps=1e-5)
        if training and (q + "1.num_batches_tracked") in sd:
            sd[q + "1.num_batches_tracked"] += 1
        h = F.leaky_relu(h, 0.01)
    h = torch.flatten(h, 1)
    mu = F.linear(h, sd[p + "fc_mu.weight"], sd[p + "fc_mu.bias"])
    var = F.softplus(F.linear(h, sd[p + "fc_var.weight"], sd[p + "fc_var.bias"])) + 1e-8
    return mu, var


def causal_masking(u, A, n_vars):
    """CausalModeling.causal_masking (ref nn.py:290-295): z_pre[b,i,:] = sum_j A[j,i] u[b,j,:] (one hop, Q3)."""
    u = u.reshape(u.shape[0], n_vars, -1)
    return torch.matmul(torch.as_tensor(A, dtype=torch.float32, device=u.device).t(), u)


def nonlinearity_add_back_noise(sd, u, z_pre, n_vars, p="causal_mask."):
    """CausalModeling.nonlinearity_add_back_noise (ref nn.py:297-312, MLP nn.py:225-240):
    z_post_i = W2_i LeakyReLU(W1_i z_pre_i + b1_i) + b2_i + u_i."""
    B = u.shape[0]
    u = u.reshape(B, n_vars, -1)
    outs = []
    for i in range(n_vars):
        q = f"{p}nonlinearities.{i}.net."
        h = F.leaky_relu(F.linear(z_pre[:, i, :], sd[q + "0.weight"], sd[q + "0.bias"]), 0.01)
        outs.append(F.linear(h, sd[q + "2.weight"], sd[q + "2.bias"]) + u[:, i, :])
    return torch.stack(outs, dim=1).reshape(B, -1)


def block_layout(cfg):
    """Static structure of UNetModel.__init__ (ref unet.py:389-491): list of (container_key, [layer kinds])."""
    mc = cfg.model_channels
    inp = [("input_blocks.0.", [("stem", cfg.in_channels, mc)])]
    chans, ch, ds = [mc], mc, 1
    idx = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_ds:
                layers.append(("attn", ch, cfg.num_heads))
            inp.append((f"input_blocks.{idx}.", layers)); idx += 1
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inp.append((f"input_blocks.{idx}.", [("down", ch, ch)])); idx += 1
            chans.append(ch)
            ds *= 2
    mid = ("middle_block.", [("res", ch, ch), ("attn", ch, cfg.num_heads), ("res", ch, ch)])
    out, idx = [], 0
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            skip = chans.pop()
            layers = [("res", ch + skip, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_ds:
                layers.append(("attn", ch, cfg.num_heads_upsample))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", ch, ch))
                ds //= 2
            out.append((f"output_blocks.{idx}.", layers)); idx += 1
    return inp, mid, out


def _run_container(sd, cfg, prefix, layers, h, emb):
    for j, (kind, a, b) in enumerate(layers):
        p = f"{prefix}{j}."
        if kind == "stem":
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
        elif kind == "res":
            h = resblock(sd, p, h, emb, cfg.use_scale_shift_norm)
        elif kind == "attn":
            h = attention_block(sd, p, h, b)
        elif kind == "down":
            h = downsample(sd, p, h)
        elif kind == "up":
            h = upsample(sd, p, h)
    return h


def embedding_trunk(sd, cfg, t, y=None, c=None):
    """time_embed(timestep_embedding(t)) + label_emb(y) + c_emb(c)  (ref unet.py:545-554)."""
    e = timestep_embedding(t, cfg.model_channels)
    e = F.linear(silu(F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                 sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if cfg.num_classes is not None:
        assert y is not None and y.shape == (t.shape[0],)
        e = e + F.embedding(y, sd["label_emb.weight"])
    if cfg.c_dim is not None:
        e = e + F.linear(silu(F.linear(c, sd["c_emb.0.weight"], sd["c_emb.0.bias"])),
                         sd["c_emb.2.weight"], sd["c_emb.2.bias"])
    return e


def torso(sd, cfg, x, emb):
    """input_blocks -> middle_block -> output_blocks(cat skip) -> out  (ref unet.py:622-632)."""
    inp, mid, out = block_layout(cfg)
    hs, h = [], x
    for prefix, layers in inp:
        h = _run_container(sd, cfg, prefix, layers, h, emb)
        hs.append(h)
    h = _run_container(sd, cfg, mid[0], mid[1], h, emb)
    for prefix, layers in out:
        h = _run_container(sd, cfg, prefix, layers, torch.cat([h, hs.pop()], dim=1), emb)
    h = silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"]))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def unet_forward(sd, cfg, x, timesteps, y=None, c=None, x_start=None, z=None, xi=None, mask_draw=None,
                 training=True):
    """UNetModel.forward (ref unet.py:525-632) -> (eps, mu, var, z_post, mask).

    RNG is injected: `xi` ~ N(0,1) [B, rep_dim] replaces the CPU-generator draw of reparameterize
    (ref nn.py:460-467) and `mask_draw` in {0,1} [B] replaces th.bernoulli(1 - drop_prob) (ref unet.py:601);
    when omitted they are drawn from torch's CPU generator in the reference's order (xi first, then mask).
    """
    emb = embedding_trunk(sd, cfg, timesteps, y, c)
    mu = var = z_post = mask = None
    if cfg.rep_dim is not None:
        if z is None:
            mu, var = encoder_encode(sd, cfg, x_start, training=training)
            if cfg.causal_modeling:
                z_pre = causal_masking(mu, cfg.A, cfg.n_vars)
                z_post = nonlinearity_add_back_noise(sd, mu, z_pre, cfg.n_vars)
                base = z_post
            else:
                base = mu
            if xi is None:
                xi = torch.randn(base.size()).to(base.device)
            z = base + ((var * 0.001) ** 0.5) * xi          # reparameterize(m, v*0.001), ref unet.py:592,594
            if cfg.masking:
                if mask_draw is None:
                    mask_draw = torch.bernoulli(torch.zeros(z.shape[0]) + (1 - cfg.drop_prob)).to(z.device)
                m = mask_draw.float()[:, None]
                z = (z * m).float()
                z_post = (z_post * m).float()                 # ref unet.py:608-611 (requires causal_modeling)
                mask = mask_draw.float()
        emb = emb + F.linear(z, sd["up_emb.weight"], sd["up_emb.bias"])
    return torso(sd, cfg, x, emb), mu, var, z_post, mask


# ----------------------------------------------------------------------------- parameter factory
def param_shapes(cfg):
    """Ordered (name, shape, kind) of every parameter/buffer, in the reference's registration order
    (ref unet.py:302-499; state_dict wire format of SURVEY.md Appendix F). kind in
    {'w','b','gn_w','gn_b','zero_w','zero_b','bn_w','bn_b','bn_rm','bn_rv','bn_n','emb'}."""
    mc, ted = cfg.model_channels, cfg.model_channels * 4
    out = [("time_embed.0.weight", (ted, mc), "w"), ("time_embed.0.bias", (ted,), "b"),
           ("time_embed.2.weight", (ted, ted), "w"), ("time_embed.2.bias", (ted,), "b")]
    if cfg.num_classes is not None:
        out.append(("label_emb.weight", (cfg.num_classes, ted), "emb"))
    if cfg.c_dim is not None:
        out += [("c_emb.0.weight", (256, cfg.c_dim), "w"), ("c_emb.0.bias", (256,), "b"),
                ("c_emb.2.weight", (ted, 256), "w"), ("c_emb.2.bias", (ted,), "b")]
    if cfg.rep_dim is not None:
        cin = cfg.in_channels
        for k, hd in enumerate(cfg.encoder_dims):
            q = f"rep_emb.encoder.{k}."
            out += [(q + "0.weight", (hd, cin, 3, 3), "w"), (q + "0.bias", (hd,), "b"),
                    (q + "1.weight", (hd,), "bn_w"), (q + "1.bias", (hd,), "bn_b"),
                    (q + "1.running_mean", (hd,), "bn_rm"), (q + "1.running_var", (hd,), "bn_rv"),
                    (q + "1.num_batches_tracked", (), "bn_n")]
            cin = hd
        fin = cfg.encoder_dims[-1] * 4
        out += [("rep_emb.fc_mu.weight", (cfg.rep_dim, fin), "w"), ("rep_emb.fc_mu.bias", (cfg.rep_dim,), "b"),
                ("rep_emb.fc_var.weight", (cfg.rep_dim, fin), "w"), ("rep_emb.fc_var.bias", (cfg.rep_dim,), "b"),
                ("up_emb.weight", (ted, cfg.rep_dim), "w"), ("up_emb.bias", (ted,), "b")]
    if cfg.causal_modeling:
        d = cfg.rep_dim // cfg.n_vars
        for i in range(cfg.n_vars):
            q = f"causal_mask.nonlinearities.{i}.net."
            out += [(q + "0.weight", (cfg.rep_dim, d), "w"), (q + "0.bias", (cfg.rep_dim,), "b"),
                    (q + "2.weight", (d, cfg.rep_dim), "w"), (q + "2.bias", (d,), "b")]

    def res(p, cin, cout):
        e = 2 * cout if cfg.use_scale_shift_norm else cout
        r = [(p + "in_layers.0.weight", (cin,), "gn_w"), (p + "in_layers.0.bias", (cin,), "gn_b"),
             (p + "in_layers.2.weight", (cout, cin, 3, 3), "w"), (p + "in_layers.2.bias", (cout,), "b"),
             (p + "emb_layers.1.weight", (e, ted), "w"), (p + "emb_layers.1.bias", (e,), "b"),
             (p + "out_layers.0.weight", (cout,), "gn_w"), (p + "out_layers.0.bias", (cout,), "gn_b"),
             (p + "out_layers.3.weight", (cout, cout, 3, 3), "zero_w"), (p + "out_layers.3.bias", (cout,), "zero_b")]
        if cin != cout:
            r += [(p + "skip_connection.weight", (cout, cin, 1, 1), "w"), (p + "skip_connection.bias", (cout,), "b")]
        return r

    def attn(p, ch):
        return [(p + "norm.weight", (ch,), "gn_w"), (p + "norm.bias", (ch,), "gn_b"),
                (p + "qkv.weight", (3 * ch, ch, 1), "w"), (p + "qkv.bias", (3 * ch,), "b"),
                (p + "proj_out.weight", (ch, ch, 1), "zero_w"), (p + "proj_out.bias", (ch,), "zero_b")]

    inp, mid, outb = block_layout(cfg)
    for prefix, layers in inp + [mid] + outb:
        for j, (kind, a, b) in enumerate(layers):
            p = f"{prefix}{j}."
            if kind == "stem":
                out += [(p + "weight", (b, a, 3, 3), "w"), (p + "bias", (b,), "b")]
            elif kind == "res":
                out += res(p, a, b)
            elif kind == "attn":
                out += attn(p, a)
            elif kind == "down":
                out += [(p + "op.weight", (a, a, 3, 3), "w"), (p + "op.bias", (a,), "b")]
            elif kind == "up":
                out += [(p + "conv.weight", (a, a, 3, 3), "w"), (p + "conv.bias", (a,), "b")]
    out += [("out.0.weight", (mc,), "gn_w"), ("out.0.bias", (mc,), "gn_b"),
            ("out.2.weight", (cfg.out_channels, mc, 3, 3), "zero_w"), ("out.2.bias", (cfg.out_channels,), "zero_b")]
    return out


def seeded_state_dict(cfg, seed=0, dezero_seed=1):
    """Deterministic synthetic weights for parity runs, independent of torch's module-init RNG order:
    weights ~ N(0, 1/fan_in), biases ~ 0.02 N(0,1), norm gains 1 + 0.1 N, norm biases 0.1 N.  The reference's
    zero_module tensors (Q5) are de-zeroed the same way when `dezero_seed` is not None (oracle patch 4), else 0."""
    g = torch.Generator().manual_seed(seed)
    gz = torch.Generator().manual_seed(dezero_seed) if dezero_seed is not None else None
    sd = {}
    for name, shape, kind in param_shapes(cfg):
        if kind in ("w", "zero_w", "emb"):
            gen = gz if kind == "zero_w" else g
            if gen is None:
                sd[name] = torch.zeros(shape)
                continue
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            std = 1.0 if kind == "emb" else fan_in ** -0.5
            sd[name] = torch.randn(shape, generator=gen) * std
        elif kind in ("b", "zero_b"):
            gen = gz if kind == "zero_b" else g
            sd[name] = torch.zeros(shape) if gen is None else 0.02 * torch.randn(shape, generator=gen)
        elif kind in ("gn_w", "bn_w"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind in ("gn_b", "bn_b"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_rm":
            sd[name] = torch.zeros(shape)
        elif kind == "bn_rv":
            sd[name] = torch.ones(shape)
        elif kind == "bn_n":
            sd[name] = torch.tensor(0, dtype=torch.long)
    return sd


def trainable_names(cfg):
    return [n for n, _, k in param_shapes(cfg) if not k.startswith("bn_r") and k != "bn_n"]

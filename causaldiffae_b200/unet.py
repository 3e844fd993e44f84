"""CausalDiffAE UNet behind the reference API (ref improved_diffusion/unet.py).

Module tree, constructor arguments, `forward` signature / 5-tuple return and every state_dict key and shape are the
reference's, so reference checkpoints load and the training / counterfactual scripts run
unchanged.  The modules are parameter containers: the torso (stem .. out conv) is executed by `engine.Engine`, a
static plan of hand-written sm_100a kernels over NHWC bf16 activations; the [B, rep_dim]-sized representation path
stays in small fp32 device ops.  Documented deviations from the shipped reference: the encoder depth
follows the image size, the DAG adjacency `A` is injectable, and the classifier-free mask follows `rep_dim`.
"""
from abc import abstractmethod

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from .nn import (SiLU, conv_nd, linear, avg_pool_nd, zero_module, normalization, timestep_embedding,  # noqa: F401
                 reparameterize, GaussianConvEncoder, CausalModeling, encoder_hidden_dims)

DAGS = {
    "morphomnist": [[0, 1], [0, 0]],
    "circuit": [[0, 1, 1, 1], [0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 0, 0]],
    "pendulum": [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]],
}


class TimestepBlock(nn.Module):
    @abstractmethod
    def forward(self, x, emb):
        """apply the module to `x` given `emb` timestep embeddings"""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """ref unet.py:36-48"""

    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, TimestepBlock) else layer(x)
        return x


class Upsample(nn.Module):
    """ref unet.py:51-79: nearest x2 + 3x3 conv"""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        assert dims == 2 and use_conv, "only the 2-D learned-conv resampling of the reference configs is built"
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        self.conv = conv_nd(dims, channels, channels, 3, padding=1)

    def forward(self, x):
        from .engine import run_layer
        assert x.shape[1] == self.channels
        return run_layer(self, x)


class Downsample(nn.Module):
    """ref unet.py:82-105: 3x3 stride-2 conv"""

    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        assert dims == 2 and use_conv, "only the 2-D learned-conv resampling of the reference configs is built"
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        self.op = conv_nd(dims, channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        from .engine import run_layer
        assert x.shape[1] == self.channels
        return run_layer(self, x)


class ResBlock(TimestepBlock):
    """ref unet.py:108-198"""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False):
        super().__init__()
        assert dims == 2
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, use_scale_shift_norm
        self.in_layers = nn.Sequential(normalization(channels), SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(SiLU(), linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm
                                                       else self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            raise NotImplementedError("3x3 skip convolutions are unused by the reference configs")
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x, emb):
        from .engine import run_layer
        return run_layer(self, x, emb)


class QKVAttention(nn.Module):
    """ref unet.py:234-253: qkv [N, 3*ch, T] -> [N, ch, T] (one fused kernel)"""

    def forward(self, qkv):
        from . import ops
        n, c3, t = qkv.shape
        x = qkv.permute(0, 2, 1).contiguous().to(th.bfloat16).reshape(n, t, c3)
        out, _ = ops.attn_fwd(x, 1)
        return out.permute(0, 2, 1).to(qkv.dtype)


class AttentionBlock(nn.Module):
    """ref unet.py:201-231"""

    def __init__(self, channels, num_heads=1, use_checkpoint=False):
        super().__init__()
        self.channels, self.num_heads, self.use_checkpoint = channels, num_heads, use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttention()
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def forward(self, x):
        from .engine import run_layer
        return run_layer(self, x)


class UNetModel(nn.Module):
    """ref unet.py:279-632"""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, c_dim=None, rep_dim=None,
                 causal_modeling=False, flow_based=False, use_checkpoint=False, num_heads=1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, masking=False, n_vars=4, image_size=None, A=None):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if flow_based:
            raise NotImplementedError("flow_based is never enabled by the reference scripts (hard-coded dim=2,k=256)")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, attention_resolutions
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.num_classes, self.c_dim, self.rep_dim = num_classes, c_dim, rep_dim
        self.use_checkpoint, self.num_heads, self.num_heads_upsample = use_checkpoint, num_heads, num_heads_upsample
        self.causal_modeling, self.flow_based, self.masking = causal_modeling, flow_based, masking
        self.drop_prob = 0.5
        self.n_vars = n_vars
        self.image_size = image_size
        # adjacency: default = the graph the reference hard-codes in forward (unet.py:571-575); injectable (Q2)
        self.A = [list(r) for r in (A if A is not None else (DAGS["morphomnist"] if n_vars == 2 else DAGS["circuit"]))]

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), SiLU(),
                                        linear(time_embed_dim, time_embed_dim))
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)
        if self.c_dim is not None:
            self.c_emb = nn.Sequential(linear(self.c_dim, 256), SiLU(), linear(256, time_embed_dim))
        if self.rep_dim is not None:
            dims_enc = encoder_hidden_dims(image_size, n_vars) if image_size is not None else None
            self.rep_emb = GaussianConvEncoder(in_channels=in_channels, latent_dim=self.rep_dim, hidden_dims=dims_enc,
                                               num_vars=n_vars if image_size is not None else 4)
            self.up_emb = nn.Linear(self.rep_dim, time_embed_dim)
        if self.causal_modeling:
            self.causal_mask = CausalModeling(latent_dim=rep_dim, num_var=self.n_vars, learn=False)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        input_block_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, dims=dims,
                                   use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm),
            AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads),
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + input_block_chans.pop(), time_embed_dim, dropout,
                                   out_channels=model_channels * mult, dims=dims, use_checkpoint=use_checkpoint,
                                   use_scale_shift_norm=use_scale_shift_norm)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))
        self._engine = None

    # ------------------------------------------------------------------ precision API of the reference
    def convert_to_fp16(self):
        """ref unet.py:501-507. The engine always computes the torso on bf16 tensor cores from fp32 master weights
        (no loss scaling needed), so this only records the request."""
        self._fp16_requested = True

    def convert_to_fp32(self):
        self._fp16_requested = False

    @property
    def inner_dtype(self):
        return next(self.input_blocks.parameters()).dtype

    @property
    def engine(self):
        from .engine import Engine
        if self._engine is None:
            object.__setattr__(self, "_engine", Engine(self))
        return self._engine

    def _apply(self, fn, *a, **k):
        # parameters are re-homed by .to()/.cuda(): the flat arena of a previous engine is stale
        object.__setattr__(self, "_engine", None)
        return super()._apply(fn, *a, **k)

    def zero_grad(self, set_to_none=False):
        """Gradients are views of the engine's flat gradient arena (the fused optimizer and the NCCL all-reduce read that
        buffer): zero it in place; never drop the views (torch's default set_to_none=True would detach `.grad` from it)."""
        if self._engine is not None:
            self._engine.grad_arena.zero_()
            self._engine.rebind_grads(discard=True)
            return
        super().zero_grad(set_to_none=set_to_none)

    # ------------------------------------------------------------------ forward
    def embed(self, timesteps, y=None, c=None):
        """time_embed(timestep_embedding(t)) (+ label_emb(y)) (+ c_emb(c))   (ref unet.py:545-554)"""
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        if self.num_classes is not None:
            assert y.shape == (timesteps.shape[0],)
            emb = emb + self.label_emb(y)
        if self.c_dim is not None:
            emb = emb + self.c_emb(c)
        return emb

    def _adjacency(self, A, device):
        """device copy of the DAG adjacency, uploaded once per device (the reference re-uploads it every forward)"""
        if A is not None:
            return th.as_tensor(A, dtype=th.float32, device=device)
        cache = self.__dict__.setdefault("_A_dev", {})
        t = cache.get(str(device))
        if t is None:
            t = cache[str(device)] = th.tensor(self.A, dtype=th.float32, device=device)
        return t

    def forward(self, x, timesteps, y=None, c=None, x_start=None, z=None, A=None, mask=None, _tmap=None, _tscale=0.0):
        """ref unet.py:525-632 -> (eps, mu, var, z_post, mask).  Every stage is a launch sequence of hand-written kernels:
        conv encoder (rep.EncoderRunner), DAG layer (csrc/small.cu), reparameterisation + keep mask (latent kernels),
        embedding trunk + all FiLM projections (rep.TrunkRunner), torso (engine.Plan)."""
        from .rep import _FilmFn, _LatentFn, anchor
        from .nn import _randn_like_ref
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        mu = var = z_post = None
        mask = None
        if self.rep_dim is not None:
            if z is None:
                mu, var = self.rep_emb.encode(x_start)
                zp = self.causal_mask(mu, self._adjacency(A, mu.device)) if self.causal_modeling else mu
                xi = _randn_like_ref(mu)                       # reference order of CPU-generator draws: xi, then the mask
                keep = None
                if self.masking:
                    keep = th.bernoulli(th.zeros(mu.shape[0]) + (1 - self.drop_prob)).to(mu.device)
                z, zp_m = _LatentFn.apply(var, zp, xi, keep, 0.001, self.n_vars)
                if self.causal_modeling:
                    z_post = zp_m
                mask = keep
        film = _FilmFn.apply(self, anchor(x.device), timesteps, y, c, z, _tmap, _tscale, th.is_grad_enabled())
        eps = self.engine.torso_film(x, film)
        return eps, mu, var, z_post, mask

"""Kernel engine for the UNet torso: a static execution plan of libcdae launches over NHWC bf16 activations.

Design (B200-first, not a translation of the eager reference):
  * all parameters live in ONE flat fp32 arena (and their gradients in a second one): the optimizer is a single
    fused kernel over the arena, the DDP all-reduce is one contiguous NCCL call, EMA is a flat copy;
    conv weights are stored OHWI (channels-last) so that the bf16 operand copies are a cast (+pad) and the
    tensor-core weight-gradient kernel writes coalesced rows;
  * the torso (stem -> input/middle/output blocks -> out conv, ref unet.py:622-632) is compiled per batch size into a
    list of kernel launches with preallocated buffers and prebuilt descriptors; forward and backward are each
    replayed as one CUDA graph;
  * concat skip connections are never materialised: GroupNorm and the implicit GEMM read two sources (split-K);
  * the ResBlock 1x1 skip conv is extra K segments of the second 3x3 implicit GEMM; bias/residual adds are epilogues.
The per-ResBlock FiLM vectors `emb_layers(SiLU(emb))` (ref unet.py:148-154,186-192) are ONE batched GEMM because the
22 emb_layers weights are adjacent in the arena.
"""
import ctypes as C
import os

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from ._lib import PackEntry

bf16 = th.bfloat16
USE_GRAPHS = os.environ.get("CDAE_GRAPHS", "1") != "0"
# GroupNorm statistics from the producing conv's epilogue + one streaming normalise pass (0: reduce inside the GN kernel)
FUSED_GN_STATS = os.environ.get("CDAE_FUSED_GN_STATS", "1") != "0"
# GroupNorm backward statistics from the epilogue of the data-gradient conv that produces the norm's output gradient + one
# streaming apply pass (cdae_igemm_desc.gnb_*, cdae_gn_bwd_apply) instead of the resident cluster kernel that reduces and
# applies in one launch.  Built, parity-tested (tests/test_kernels_gpu.py::test_groupnorm_backward_from_dgrad_epilogue, the
# per-layer and model-level tests pass with it on) and MEASURED on B200 (profiles/r2_gnb_bench.log, r2_ncu_igemm3_gnb_*):
# the norm's own backward gets faster (128 ch @ 64x64: 76 -> 57 us, 0.40 -> 0.54 of the HBM peak) but the conv pays more than
# that in its epilogue (67 -> 124 us): silu' + the register butterfly add ~1400 instructions per thread and slab to the ONE
# epilogue warp per scheduler, which cannot hide its own latencies (tensor pipe 53 % -> 28 %).  Whole step 19.25 -> 20.3 ms.
# Off by default until the epilogue runs two warps per scheduler.
FUSED_GN_BWD = os.environ.get("CDAE_FUSED_GN_BWD", "0") != "0"
# Weight-gradient kernels on a second stream: nothing in the backward chain (data gradients, GroupNorm / attention backward)
# reads a weight gradient, so every wgrad launch only has to wait for the kernel that produced its dY and has to be done by
# the end of the backward.  In the captured graph this becomes a side branch: its CTAs fill the tail waves and launch gaps of
# the chain (both kinds of kernel occupy a whole SM per CTA, so this is interleaving, not co-residency).
# stride-2 data gradients as four parity-class launches (0: one full-resolution conv over a zero-inserted dy)
S2_DGRAD_CLASSES = os.environ.get("CDAE_S2_DGRAD_CLASSES", "1") != "0"
# inference: GroupNorm(+FiLM)+SiLU applied by the consuming 3x3 conv while it loads its operand (no activated tensor)
GN_ON_LOAD = os.environ.get("CDAE_GN_ON_LOAD", "1") != "0"
# inference: nearest x2 upsampling expanded by the consuming conv from a low-resolution halo tile (no 4x tensor)
UP_ON_LOAD = os.environ.get("CDAE_UP_ON_LOAD", "1") != "0"
WGRAD_SIDE_STREAM = os.environ.get("CDAE_WGRAD_SIDE_STREAM", "1") != "0"


def _round_up(v, m):
    return (v + m - 1) // m * m


class _ConvW:
    """bf16 operand copies of one conv/linear weight inside the bf16 arena."""

    def __init__(self, param, bias, cout, cin, taps, cin_pad=None, cout_pad=None):
        self.param, self.bias = param, bias
        self.cout, self.cin, self.taps = cout, cin, taps
        self.cin_pad = cin_pad or _round_up(cin, 64)
        self.cout_pad = cout_pad or (_round_up(cout, 64) if cout >= 64 else 16)
        self.fwd = None      # bf16 [cout_pad, K]
        self.tr = None       # bf16 [cin_pad, taps*cout_tr_pad]
        self.cout_tr_pad = _round_up(cout, 64)


class T:
    """A planned activation: bf16 NHWC buffer + lazily allocated gradient buffer."""

    def __init__(self, plan, shape, name=""):
        self.plan, self.shape, self.name = plan, tuple(shape), name
        self.t = plan.alloc(shape)
        self.g = None
        self.g_written = False
        self.stats = None       # fp32 [B, C, 2] channel sums written by the producing conv's epilogue (GroupNorm input)
        self.gnb = None         # set on a GroupNorm OUTPUT: dict(x0, x1, ab, ws, silu) for the fused backward statistics

    def grad(self):
        if self.g is None:
            self.g = self.plan.alloc(self.shape)
        return self.g

    def grad_acc(self):
        """False for the first producer of this gradient (write), True for later ones (accumulate)."""
        acc = self.g_written
        self.g_written = True
        return acc


class Plan:
    def __init__(self, engine, B, train):
        self.e, self.B, self.train = engine, B, train
        self.dev = engine.device
        self.fwd, self.bwd = [], []        # lists of zero-arg callables; bwd is executed in reverse build order
        self.bufs = []
        self.fwd_graph = self.bwd_graph = None
        self.runs = 0
        self.generation, self.pending = 0, False     # which forward the saved activations belong to / awaits backward
        # nn.Dropout of the ResBlocks (ref unet.py:157): counter-based masks, {seed, base} on the device; the base moves on at
        # the start of every forward, so the backward of the same step regenerates the same masks; p = 0 in eval mode
        self.drop_state = th.tensor([int(th.initial_seed()) & 0x7fffffffffffffff, 0], device=self.dev, dtype=th.int64)
        self.drop_p = th.zeros(1, device=self.dev)
        self.drop_p_host, self.drop_groups = 0.0, 0
        self.n_fwd_launch = self.n_bwd_launch = 0
        # per-(image, channel) GroupNorm sums accumulated by conv epilogues: carved from a few big chunks that the
        # first node of the forward graph zeroes
        self.stat_chunks, self._stat_used = [], 0
        self.bws_chunks, self._bws_used = [], 0
        self.add_fwd(self._zero_stats, 0)

    def alloc(self, shape, dtype=bf16):
        t = th.empty(shape, device=self.dev, dtype=dtype)
        self.bufs.append(t)
        return t

    def alloc_stats(self, B, C):
        n = B * C * 2
        if not self.stat_chunks or self._stat_used + n > self.stat_chunks[-1].numel():
            self.stat_chunks.append(th.zeros(max(n, 1 << 21), device=self.dev, dtype=th.float32))
            self._stat_used = 0
            self.n_fwd_launch += 1
        v = self.stat_chunks[-1][self._stat_used:self._stat_used + n].view(B, C, 2)
        self._stat_used += (n + 3) // 4 * 4
        return v

    def _zero_stats(self):
        for c in self.stat_chunks:
            ops.zero_(c)
        if self.drop_groups:
            ops.step_tick(self.drop_state[1:], None, self.drop_groups)

    def set_dropout(self, p):
        if p != self.drop_p_host:
            self.drop_p.fill_(p)
            self.drop_p_host = p

    def alloc_bwd_ws(self, B, C):
        """zeroed-at-backward-start fp32 [B, C, 2] {sum du, sum du*x} accumulator of one fused GroupNorm backward"""
        n = B * C * 2
        if not self.bws_chunks or self._bws_used + n > self.bws_chunks[-1].numel():
            self.bws_chunks.append(th.zeros(max(n, 1 << 21), device=self.dev, dtype=th.float32))
            self._bws_used = 0
        v = self.bws_chunks[-1][self._bws_used:self._bws_used + n].view(B, C, 2)
        self._bws_used += (n + 3) // 4 * 4
        return v

    def _zero_bwd_ws(self):
        for c in self.bws_chunks:
            ops.zero_(c)

    def add_fwd(self, fn, n=1):
        self.fwd.append(fn); self.n_fwd_launch += n

    def add_bwd(self, fns):
        """fns: callables of ONE layer in execution order; layers are replayed last-to-first."""
        self.bwd.append(list(fns)); self.n_bwd_launch += len(fns)

    # ---- execution
    def _run_fwd_eager(self):
        for f in self.fwd:
            f()

    def _run_bwd_eager(self):
        if not WGRAD_SIDE_STREAM:
            for layer in reversed(self.bwd):
                for f in layer:
                    f()
            return
        main = th.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = th.cuda.Stream()
        side = self._side
        used = False
        for layer in reversed(self.bwd):
            for f in layer:
                if getattr(f, "side", False):
                    side.wait_stream(main)           # after everything issued so far: in particular the producer of its dY
                    with th.cuda.stream(side):
                        f()
                    used = True
                else:
                    f()
        if used:
            main.wait_stream(side)                   # the optimizer (and the next step's zeroing) come after every wgrad

    def forward(self):
        if USE_GRAPHS and self.runs >= 1:
            if self.fwd_graph is None:
                self.fwd_graph = _capture(self._run_fwd_eager)
            self.fwd_graph.replay()
        else:
            self._run_fwd_eager()
        self.runs += 1

    def backward(self):
        if USE_GRAPHS and self.runs >= 2:
            if self.bwd_graph is None:
                self.bwd_graph = _capture(self._run_bwd_eager)
            self.bwd_graph.replay()
        else:
            self._run_bwd_eager()


def _side(fn):
    """mark a backward launch as independent of the chain (see WGRAD_SIDE_STREAM)"""
    fn.side = True
    return fn


def _capture(fn):
    g = th.cuda.CUDAGraph()
    th.cuda.synchronize()
    s = th.cuda.Stream()
    s.wait_stream(th.cuda.current_stream())
    with th.cuda.stream(s):
        with th.cuda.graph(g, stream=s):
            fn()
    th.cuda.current_stream().wait_stream(s)
    return g


class Engine:
    def __init__(self, model):
        self.model = model
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise _lib.CdaeError("the UNet engine needs the model on a CUDA (sm_100a) device; there is no CPU path")
        self.device = p0.device
        _lib.lib()
        self.plans = {}
        self._layer_plans = {}
        self._build_arena()
        self._build_weight_pack()
        self._packed_version = None
        self.dirty = True
        for mod in model.modules():
            object.__setattr__(mod, "_cdae_root", model)

    # ------------------------------------------------------------------ flat parameter arena
    def _torso_conv_params(self):
        m = self.model
        out = set()
        from .unet import ResBlock, AttentionBlock, Upsample, Downsample
        for mod in list(m.input_blocks.modules()) + list(m.middle_block.modules()) + list(m.output_blocks.modules()):
            if isinstance(mod, (nn.Conv2d, nn.Conv1d)):
                out.add(id(mod.weight))
        out.add(id(m.out[2].weight))
        return out

    def resblocks(self):
        from .unet import ResBlock
        m = self.model
        return [mod for blk in list(m.input_blocks) + [m.middle_block] + list(m.output_blocks) for mod in blk
                if isinstance(mod, ResBlock)]

    def _build_arena(self):
        m = self.model
        ohwi = self._torso_conv_params()
        rbs = self.resblocks()
        order, seen = [], set()
        for rb in rbs:
            order.append(rb.emb_layers[1].weight)
        for rb in rbs:
            order.append(rb.emb_layers[1].bias)
        seen = {id(p) for p in order}
        # the parameters of the representation path (embedding trunk, conv encoder, DAG layer) go LAST: their gradients are the
        # last ones a backward pass produces, so [0, early_end) - 99 % of the arena - is final as soon as the FiLM projection's
        # backward has run and its all-reduce can start beside the rest of the backward (TrainLoop._forward_backward_fused)
        late = set()
        for name in ("time_embed", "label_emb", "c_emb", "rep_emb", "up_emb", "causal_mask"):
            mod = getattr(m, name, None)
            if isinstance(mod, nn.Module):
                late |= {id(p) for p in mod.parameters()}
        for want_late in (False, True):
            for p in m.parameters():
                if id(p) not in seen and (id(p) in late) == want_late:
                    order.append(p); seen.add(id(p))
        offs, off = {}, 0
        self.early_end = None
        for p in order:
            if id(p) in late and self.early_end is None:
                self.early_end = off
            offs[id(p)] = off
            off += _round_up(p.numel(), 4)
        if self.early_end is None:
            self.early_end = off
        self.n_params = off
        self.arena = th.zeros(off, device=self.device, dtype=th.float32)
        self.grad_arena = th.zeros(off, device=self.device, dtype=th.float32)
        self.param_offsets = offs
        self._views = []     # (param, offset, numel, ohwi?) for state export from other arenas (EMA)
        with th.no_grad():
            for p in order:
                o, n = offs[id(p)], p.numel()
                is_ohwi = id(p) in ohwi and p.dim() == 4
                flat, gflat = self.arena[o:o + n], self.grad_arena[o:o + n]
                if is_ohwi:
                    O, I, H, W = p.shape
                    flat.view(O, H, W, I).copy_(p.detach().permute(0, 2, 3, 1))
                    p.data = flat.view(O, H, W, I).permute(0, 3, 1, 2)
                    p.grad = gflat.view(O, H, W, I).permute(0, 3, 1, 2)
                else:
                    flat.view(p.shape).copy_(p.detach())
                    p.data = flat.view(p.shape)
                    p.grad = gflat.view(p.shape)
                self._views.append((p, o, n, is_ohwi))
        ted = m.model_channels * 4
        self.film_width = sum(rb.emb_layers[1].weight.shape[0] for rb in rbs)
        self.film_w = nn.Parameter(self.arena[:self.film_width * ted].view(self.film_width, ted))
        self.film_w.grad = self.grad_arena[:self.film_width * ted].view(self.film_width, ted)
        b0 = offs[id(rbs[0].emb_layers[1].bias)]
        self.film_b = nn.Parameter(self.arena[b0:b0 + self.film_width])
        self.film_b.grad = self.grad_arena[b0:b0 + self.film_width]
        self.film_off, o = {}, 0
        for rb in rbs:
            self.film_off[id(rb)] = o
            o += rb.emb_layers[1].weight.shape[0]

    def owns(self, p):
        a0 = self.arena.data_ptr()
        return a0 <= p.data_ptr() < a0 + self.arena.numel() * 4

    def owns_grad(self, g):
        a0 = self.grad_arena.data_ptr()
        return a0 <= g.data_ptr() < a0 + self.grad_arena.numel() * 4

    def rebind_grads(self, discard=False):
        """re-attach every Parameter's .grad to its view of the gradient arena (after zero_grad(set_to_none=True) or any
        code that replaced .grad by a fresh tensor); a stray gradient is folded into the arena unless `discard`"""
        for p, o, n, is_ohwi in self._views:
            if p.grad is not None and self.owns_grad(p.grad):
                continue
            stray = p.grad
            gflat = self.grad_arena[o:o + n]
            if is_ohwi:
                O, I, H, W = p.shape
                p.grad = gflat.view(O, H, W, I).permute(0, 3, 1, 2)
            else:
                p.grad = gflat.view(p.shape)
            if stray is not None and not discard:
                with th.no_grad():
                    p.grad.add_(stray)

    def grad_of(self, p):
        o, n = self.param_offsets[id(p)], p.numel()
        return self.grad_arena[o:o + n]

    def export_state(self, flat):
        """{name: tensor} views of another flat arena (e.g. EMA) with the model's names/shapes."""
        by_id = {id(p): (o, n, oh) for p, o, n, oh in self._views}
        out = {}
        for name, p in self.model.named_parameters():
            o, n, oh = by_id[id(p)]
            if oh:
                O, I, H, W = p.shape
                out[name] = flat[o:o + n].view(O, H, W, I).permute(0, 3, 1, 2)
            else:
                out[name] = flat[o:o + n].view(p.shape)
        return out

    # ------------------------------------------------------------------ bf16 operand copies
    def _build_weight_pack(self):
        from .unet import ResBlock, AttentionBlock, Upsample, Downsample
        m = self.model
        self.convs = {}
        entries, off = [], 0

        def reg(key, w, b, cout, cin, taps, cin_pad=None, cout_pad=None, fuse_skip=None):
            nonlocal off
            cw = _ConvW(w, b, cout, cin, taps, cin_pad, cout_pad)
            K = taps * cw.cin_pad
            Ktot = K + (fuse_skip.cin_pad if fuse_skip is not None else 0)
            cw.K, cw.Ktot = K, Ktot
            cw.fwd_off, cw.fwd_shape = off, (cw.cout_pad, Ktot)
            off += _round_up(cw.cout_pad * Ktot, 64)
            cw.tr_off, cw.tr_shape = off, (cw.cin_pad, taps * cw.cout_tr_pad)
            off += _round_up(cw.cin_pad * taps * cw.cout_tr_pad, 64)
            e = PackEntry(self.param_offsets[id(w)], cw.fwd_off, cw.tr_off, cout, cin, taps, cw.cout_pad, cw.cin_pad,
                          Ktot, 0, 0)
            entries.append(e)
            if fuse_skip is not None:
                # the 1x1 skip weight lives in the trailing columns of the same forward matrix
                sk = fuse_skip
                sk.fwd_in = cw
                entries.append(PackEntry(self.param_offsets[id(sk.param)], cw.fwd_off + K, -1, sk.cout, sk.cin, 1,
                                         cw.cout_pad, sk.cin_pad, Ktot, 0, 0))
            self.convs[key] = cw
            # transposed copy must use cout_tr_pad columns: separate entry when it differs from cout_pad
            if cw.cout_tr_pad != cw.cout_pad:
                e.dst_tr_off = -1
                entries.append(PackEntry(self.param_offsets[id(w)], -1, cw.tr_off, cout, cin, taps, cw.cout_tr_pad,
                                         cw.cin_pad, 0, 0, 0))
            return cw

        stem = m.input_blocks[0][0]
        reg(id(stem), stem.weight, stem.bias, stem.out_channels, stem.in_channels, 9)
        for blk in list(m.input_blocks)[1:] + [m.middle_block] + list(m.output_blocks):
            for mod in blk:
                if isinstance(mod, ResBlock):
                    c1, c2 = mod.in_layers[2], mod.out_layers[3]
                    reg(id(c1), c1.weight, c1.bias, c1.out_channels, c1.in_channels, 9)
                    sk = None
                    if not isinstance(mod.skip_connection, nn.Identity):
                        s = mod.skip_connection
                        sk = _ConvW(s.weight, s.bias, s.out_channels, s.in_channels, 1)
                        sk.tr_off, sk.tr_shape = off, (sk.cin_pad, sk.cout_tr_pad)
                        off += _round_up(sk.cin_pad * sk.cout_tr_pad, 64)
                        entries.append(PackEntry(self.param_offsets[id(s.weight)], -1, sk.tr_off, sk.cout, sk.cin, 1,
                                                 sk.cout_tr_pad, sk.cin_pad, 0, 0, 0))
                        self.convs[id(s)] = sk
                    reg(id(c2), c2.weight, c2.bias, c2.out_channels, c2.in_channels, 9, fuse_skip=sk)
                elif isinstance(mod, AttentionBlock):
                    reg(id(mod.qkv), mod.qkv.weight, mod.qkv.bias, mod.qkv.out_channels, mod.qkv.in_channels, 1)
                    reg(id(mod.proj_out), mod.proj_out.weight, mod.proj_out.bias, mod.proj_out.out_channels,
                        mod.proj_out.in_channels, 1)
                elif isinstance(mod, Downsample):
                    reg(id(mod.op), mod.op.weight, mod.op.bias, mod.op.out_channels, mod.op.in_channels, 9)
                elif isinstance(mod, Upsample):
                    reg(id(mod.conv), mod.conv.weight, mod.conv.bias, mod.conv.out_channels, mod.conv.in_channels, 9)
        oc = m.out[2]
        reg(id(oc), oc.weight, oc.bias, oc.out_channels, oc.in_channels, 9)
        # all emb_layers weights (adjacent in the arena) as ONE [film_width, time_embed_dim] matrix: the FiLM projection and
        # its two backward GEMMs run on the tensor cores too (bf16 operands, fp32 accumulate and fp32 results)
        ted = m.model_channels * 4
        if self.film_width % 64 == 0 and ted % 64 == 0:
            fw = _ConvW(self.film_w, self.film_b, self.film_width, ted, 1)
            fw.K = fw.Ktot = ted
            fw.fwd_off, fw.fwd_shape = off, (self.film_width, ted)
            off += _round_up(self.film_width * ted, 64)
            fw.tr_off, fw.tr_shape = off, (ted, self.film_width)
            off += _round_up(ted * self.film_width, 64)
            entries.append(PackEntry(0, fw.fwd_off, fw.tr_off, self.film_width, ted, 1, self.film_width, ted, ted, 0, 0))
            self.convs["film"] = fw
        self.bf16_arena = th.zeros(off, device=self.device, dtype=bf16)
        for cw in self.convs.values():
            if getattr(cw, "fwd_off", None) is not None and cw.fwd is None and hasattr(cw, "fwd_shape"):
                n = cw.fwd_shape[0] * cw.fwd_shape[1]
                cw.fwd = self.bf16_arena[cw.fwd_off:cw.fwd_off + n].view(cw.fwd_shape)
            n = cw.tr_shape[0] * cw.tr_shape[1]
            cw.tr = self.bf16_arena[cw.tr_off:cw.tr_off + n].view(cw.tr_shape)
        tiles = 0
        for e in entries:       # prefix sum of 32x32 tiles (see pack_weights_kernel)
            e._pad = tiles
            tiles += e.taps * ((e.cout_pad + 31) // 32) * ((e.cin_pad + 31) // 32)
        self.pack_tiles = tiles
        arr = (PackEntry * len(entries))(*entries)
        host = th.frombuffer(bytearray(bytes(arr)), dtype=th.uint8)
        self.pack_entries = host.to(self.device)
        self.n_pack = len(entries)
        self.pack_max = max(max(e.cout_pad, 1) * e.taps * e.cin_pad for e in entries)

    def weights_version(self):
        """Changes whenever the master weights were written through torch: the arena's own version counter plus the
        counters of the Parameters re-homed into it (`p.data = arena.view(..)` gives each Parameter its OWN counter, so
        load_state_dict / zero_module / a stock optimizer bump those and not the arena's).  Writes that bypass torch
        (the fused optimizer kernel) set `dirty`."""
        return (self.arena._version, sum(p._version for p, _, _, _ in self._views))

    def pack(self, force=False):
        """fp32 master arena -> bf16 operand copies (one launch). Re-run whenever the arena changed."""
        ver = self.weights_version()
        if not (force or self.dirty or ver != self._packed_version):
            return
        ops.check(_lib.lib().cdae_pack_weights(self.arena.data_ptr(), self.bf16_arena.data_ptr(),
                                               self.pack_entries.data_ptr(), self.n_pack, self.pack_tiles, ops.stream()))
        self._packed_version, self.dirty = ver, False

    # ------------------------------------------------------------------ planning helpers
    def _gparam(self, p):
        """fp32 gradient view of a (non-OHWI) parameter inside the grad arena"""
        return self.grad_of(p)

    def plan_conv(self, pl, cw, srcs, out, ksize, stride=1, resid=None, skip=None, skip_srcs=None, out_mode=0,
                  need_dgrad=True, stats=False, bias_img=None, gn_ab=None):
        """out = conv_ksize(concat(srcs)) + bias [+ resid] [+ conv1x1_skip(concat(skip_srcs)) + bias_skip]
        stats: `out` feeds a GroupNorm - let the epilogue accumulate its per-(image, channel) sums (out.stats).
        gn_ab: `srcs` are the RAW inputs of a GroupNorm(+FiLM)+SiLU whose constants are in gn_ab [B, C, 2]: the conv applies
        it while loading (inference, see gn_on_load_ok)."""
        chans = [s.shape[3] for s in srcs]
        segs, K = ops.conv_segments(chans, ksize)
        all_srcs = list(srcs)
        bias2 = None
        if skip is not None:
            ssegs, _ = ops.conv_segments([s.shape[3] for s in skip_srcs], 1, wk0=K, src0=len(all_srcs))
            segs += ssegs
            all_srcs += list(skip_srcs)
            bias2 = skip.bias
        st = None
        if stats and FUSED_GN_STATS and out_mode == 0 and cw.cout % 64 == 0 and cw.cout == out.shape[3] and \
                out.shape[1] * out.shape[2] >= 32:
            st = out.stats = pl.alloc_stats(out.shape[0], cw.cout)
        gn = None
        if gn_ab is not None:
            offs, o = [], 0
            for c in chans:
                offs.append(o); o += c
            gn = (gn_ab, offs + [-1] * (len(all_srcs) - len(srcs)))
        d = ops.make_igemm_desc([s.t for s in all_srcs], segs, cw.fwd, out if out_mode == 1 else out.t, cw.cout,
                                in_stride=stride, bias=cw.bias, bias2=bias2, resid=resid.t if resid is not None else None,
                                out_mode=out_mode, stats=st, bias_img=bias_img, gn=gn)
        pl.add_fwd(lambda: ops.igemm(d))
        return chans

    def plan_conv_bwd(self, pl, cw, srcs, dy, ksize, stride=1, need_dgrad=True):
        """gradients of out = conv(concat(srcs)): bias, weight, and data (into srcs[i].grad()). dy: bf16 NHWC tensor.
        A single source that is the output of a streaming GroupNorm (srcs[0].gnb) gets du = dy_src * silu'(u) instead of
        its plain gradient, plus the backward statistics of that norm, from the data-gradient epilogue."""
        fns = []
        gb, gw = self._gparam(cw.bias), self._gparam(cw.param)
        # the bias gradient (column sums of dy) rides along in the weight-gradient kernel of the first source
        fused_bias = cw.cout % 8 == 0
        if fused_bias:
            pass
        elif cw.cout % 8 == 0:
            fns.append(lambda: ops.colsum_(dy, gb, c=cw.cout))
        else:   # final eps conv: 3 real channels inside a 64-wide padded gradient
            tmp = pl.alloc((dy.shape[-1],), th.float32)
            fns.append(lambda: ops.zero_(tmp))
            fns.append(lambda: ops.colsum_(dy, tmp))
            fns.append(lambda: gb.add_(tmp[:cw.cout]))
        off = 0
        for si, s in enumerate(srcs):
            c = s.shape[3]
            real = max(0, min(c, cw.cin - off))
            wd = ops.make_wgrad_desc(dy, s.t, gw, cw.cout, c, ksize=ksize, in_stride=stride, ci_off=off, cin_real=real,
                                     dw_ld=cw.cin, dbias=gb if (fused_bias and si == 0) else None)
            fns.append(_side(lambda wd=wd: ops.wgrad(wd)))
            off += c
        if need_dgrad and stride == 2 and S2_DGRAD_CLASSES:
            # conv_transpose by output parity: dx[2m+a, 2q+b] only sees the taps kh = a+1 (mod 2), kw = b+1 (mod 2) - 1, 2, 2
            # and 4 of the 9 - at dy[m + (a+1-kh)/2, q + (b+1-kw)/2].  Four launches over the LOW-resolution dy that store
            # with pixel stride 2 (cdae_igemm_desc.sps / ooh / oow), instead of one full-resolution 9-tap conv over a
            # zero-inserted copy of dy (4x the MMA work, one more pass over HBM).  ref unet.py:97-105 backward.
            nco = dy.shape[-1]
            if nco != cw.cout_tr_pad:
                raise _lib.CdaeError("internal: dy pitch does not match the transposed weight packing")
            off = 0
            for s in srcs:
                c = s.shape[3]
                acc, g = s.grad_acc(), s.grad()
                s.gnb = None                         # a plain gradient: the norm's backward must reduce itself
                for a in range(2):
                    for b in range(2):
                        khs = [(1, 0)] if a == 0 else [(0, 1), (2, 0)]          # (tap, shift in dy)
                        kws = [(1, 0)] if b == 0 else [(0, 1), (2, 0)]
                        segs = [(0, dh, dw, 0, nco // 64, (kh * 3 + kw) * nco) for kh, dh in khs for kw, dw in kws]
                        d = ops.make_igemm_desc([dy], segs, cw.tr[off:off + c], g, c, resid=g if acc else None,
                                                sps=2, ooh=a, oow=b)
                        fns.append(lambda d=d: ops.igemm(d))
                off += c
        elif need_dgrad:
            dyz = dy
            if stride == 2:
                dyz = pl.alloc((dy.shape[0], dy.shape[1] * 2, dy.shape[2] * 2, dy.shape[3]))
                fns.append(lambda: ops.zero_insert2x(dy, out=dyz))
            nco = dy.shape[-1]
            if nco != cw.cout_tr_pad:
                raise _lib.CdaeError("internal: dy pitch does not match the transposed weight packing")
            segs, _ = ops.conv_segments([nco], ksize, transposed=True)
            off = 0
            for s in srcs:
                c = s.shape[3]
                acc, g = s.grad_acc(), s.grad()
                gnb = s.gnb if (len(srcs) == 1 and not acc and stride == 1) else None
                if s.gnb is not None and gnb is None:
                    s.gnb = None                     # this gradient is a plain dy: the norm's backward must reduce itself
                d = ops.make_igemm_desc([dyz], segs, cw.tr[off:off + c], g, c, resid=g if acc else None, gnb=gnb)
                fns.append(lambda d=d: ops.igemm(d))
                off += c
        return fns

    def plan_gn_fwd(self, pl, x0, x1, gn, out, st, film=None, film_off=0, silu=True):
        """out = [SiLU](FiLM(GroupNorm32(concat(x0, x1)))) (ref nn.py:430-437, unet.py:185-198); st = (mean, rstd) [B,32].
        Streaming kernel when every source carries channel sums from its producer, reducing kernel otherwise."""
        x1t = x1.t if x1 is not None else None
        if x0.stats is not None and (x1 is None or x1.stats is not None):
            s1 = x1.stats if x1 is not None else None
            ab = None
            Ct = out.shape[3]
            if pl.train and FUSED_GN_BWD and Ct % 64 == 0 and x0.shape[3] % 64 == 0 and out.shape[1] * out.shape[2] >= 32:
                # the data-gradient conv that will produce d(out) also produces this norm's backward statistics
                ab = pl.alloc((out.shape[0], Ct, 2), th.float32) if silu else None
                out.gnb = dict(x0=x0.t, x1=x1t, ab=ab, ws=pl.alloc_bwd_ws(out.shape[0], Ct), silu=silu)
            pl.add_fwd(lambda: ops.gn_apply_fwd(x0.t, x0.stats, gn.weight, gn.bias, x1=x1t, stats1=s1, film=film,
                                                film_off=film_off, silu=silu, out=out.t, mean=st[0], rstd=st[1], ab=ab))
        else:
            pl.add_fwd(lambda: ops.gn_fwd(x0.t, gn.weight, gn.bias, x1=x1t, film=film, film_off=film_off, silu=silu,
                                          out=out.t, mean=st[0], rstd=st[1]))

    def gn_on_load_ok(self, pl, cw, x0, x1):
        """inference only: can the 3x3 conv `cw` over [SiLU](FiLM(GroupNorm(concat(x0, x1)))) apply the norm while it loads?
        (the halo kernels: the image tiles into 8 x 16 boxes - the 64x64 ... 16x16 levels; statistics from the producers)"""
        if pl.train or not GN_ON_LOAD:
            return False
        B, H, W = x0.shape[:3]
        if x0.stats is None or (x1 is not None and x1.stats is None):
            return False
        chans_ok = x0.shape[3] % 64 == 0 and (x1 is None or x1.shape[3] % 64 == 0)
        return chans_ok and H % 16 == 0 and W % 8 == 0

    def plan_gn_constants(self, pl, x0, x1, gn, st, film=None, film_off=0):
        """the {a, b} table of [SiLU](FiLM(GroupNorm32(concat(x0, x1)))) for a conv that applies it on load"""
        Ct = x0.shape[3] + (x1.shape[3] if x1 is not None else 0)
        ab = pl.alloc((x0.shape[0], Ct, 2), th.float32)
        x1t = x1.t if x1 is not None else None
        s1 = x1.stats if x1 is not None else None
        pl.add_fwd(lambda: ops.gn_apply_fwd(x0.t, x0.stats, gn.weight, gn.bias, x1=x1t, stats1=s1, film=film, film_off=film_off,
                                            silu=True, mean=st[0], rstd=st[1], ab=ab, constants_only=True))
        return ab

    def plan_gn_bwd(self, a, x0, x1, gn, st, dx0, dx1=None, accmask=0, film=None, film_off=0, silu=True, dfilm=None, dadd=None):
        """backward of a = [SiLU](FiLM(GroupNorm32(concat(x0, x1)))): one streaming pass when the conv that produced d(a) left
        du and the statistics behind (a.gnb, see plan_conv_bwd), else the resident reduce-and-apply kernel."""
        da = a.grad()
        gw, gb = self._gparam(gn.weight), self._gparam(gn.bias)
        x1t = x1.t if x1 is not None else None
        if a.gnb is not None:
            ws = a.gnb["ws"]
            return lambda: ops.gn_bwd_apply(da, x0.t, gn.weight, gn.bias, st[0], st[1], ws, x1=x1t, film=film, film_off=film_off,
                                            dx0=dx0, dx1=dx1, accumulate_dx=accmask, dgamma=gw, dbeta=gb, dfilm=dfilm, dadd=dadd)
        return lambda: ops.gn_bwd(da, x0.t, gn.weight, gn.bias, st[0], st[1], x1=x1t, film=film, film_off=film_off, silu=silu,
                                  dx0=dx0, dx1=dx1, accumulate_dx=accmask, dgamma=gw, dbeta=gb, dfilm=dfilm, dadd=dadd)

    # ------------------------------------------------------------------ layer planners
    def plan_resblock(self, pl, rb, x0, x1, film, dfilm):
        B, H, W = x0.shape[:3]
        cin = x0.shape[3] + (x1.shape[3] if x1 is not None else 0)
        cout = rb.out_channels
        gn1, c1, gn2, c2 = rb.in_layers[0], rb.in_layers[2], rb.out_layers[0], rb.out_layers[3]
        cw1, cw2 = self.convs[id(c1)], self.convs[id(c2)]
        has_skip = not isinstance(rb.skip_connection, nn.Identity)
        sk = self.convs[id(rb.skip_connection)] if has_skip else None
        foff = self.film_off[id(rb)]
        h1, out = T(pl, (B, H, W, cout)), T(pl, (B, H, W, cout))
        a1 = a2 = None                  # the activated tensors exist only where a norm is not applied on load
        st1 = (pl.alloc((B, 32), th.float32), pl.alloc((B, 32), th.float32))
        st2 = (pl.alloc((B, 32), th.float32), pl.alloc((B, 32), th.float32))
        x1t = x1.t if x1 is not None else None
        ssn = rb.use_scale_shift_norm
        # inference at the 64x64 / 32x32 levels: the norms are applied by the consuming conv while it loads its operand -
        # no activated tensor, no extra pass over HBM (ops.make_igemm_desc(gn=...)); elsewhere one streaming pass each
        srcs_x = [x0] + ([x1] if x1 is not None else [])
        bimg = None if ssn else film[:, foff:foff + cout]      # additive conditioning rides in conv1's epilogue (unet.py:195-197)
        if self.gn_on_load_ok(pl, cw1, x0, x1):
            ab1 = self.plan_gn_constants(pl, x0, x1, gn1, st1)
            self.plan_conv(pl, cw1, srcs_x, h1, 3, stats=True, bias_img=bimg, gn_ab=ab1)
        else:
            a1 = T(pl, (B, H, W, cin))
            self.plan_gn_fwd(pl, x0, x1, gn1, a1, st1)
            self.plan_conv(pl, cw1, [a1], h1, 3, stats=True, bias_img=bimg)
        fuse2 = self.gn_on_load_ok(pl, cw2, h1, None)
        if fuse2:    # h = GN(conv1(.)) * (1 + scale) + shift  (ref unet.py:190-194), or plain GN after the additive form
            ab2 = self.plan_gn_constants(pl, h1, None, gn2, st2, film=film if ssn else None, film_off=foff if ssn else 0)
        else:
            a2 = T(pl, (B, H, W, cout))
            if ssn:
                self.plan_gn_fwd(pl, h1, None, gn2, a2, st2, film=film, film_off=foff)
            else:
                self.plan_gn_fwd(pl, h1, None, gn2, a2, st2)
        drop_off = None
        if rb.dropout and pl.train:     # nn.Dropout between SiLU and conv2 (ref unet.py:157): in place on a2, masks regenerated in bwd
            drop_off = pl.drop_groups
            pl.drop_groups += a2.t.numel() // 8
            a2.gnb = None               # d(a2) has to be masked before the norm's backward sees it
            pl.add_fwd(lambda: ops.dropout_(a2.t, pl.drop_state, drop_off, pl.drop_p))
        c2_src, c2_gn = ([h1], ab2) if fuse2 else ([a2], None)
        if has_skip:
            self.plan_conv(pl, cw2, c2_src, out, 3, skip=sk, skip_srcs=srcs_x, stats=True, gn_ab=c2_gn)
        else:
            self.plan_conv(pl, cw2, c2_src, out, 3, resid=x0, stats=True, gn_ab=c2_gn)
        if pl.train:
            def build_bwd():
                fns = []
                dout = out.grad()
                # conv2 (+ fused skip): bias, weights, data
                fns += self.plan_conv_bwd(pl, cw2, [a2], dout, 3)
                if has_skip:
                    gsk_b, gsk_w = self._gparam(sk.bias), self._gparam(sk.param)
                    off = 0
                    for s in srcs_x:
                        c = s.shape[3]
                        wd = ops.make_wgrad_desc(dout, s.t, gsk_w, cout, c, ksize=1, ci_off=off, dw_ld=cin,
                                                 dbias=gsk_b if off == 0 else None)
                        fns.append(_side(lambda wd=wd: ops.wgrad(wd)))
                        acc, g = s.grad_acc(), s.grad()
                        segs, _ = ops.conv_segments([cout], 1)
                        d = ops.make_igemm_desc([dout], segs, sk.tr[off:off + c], g, c, resid=g if acc else None)
                        fns.append(lambda d=d: ops.igemm(d))
                        off += c
                # GN2 (+FiLM) backward -> dh1
                dh1 = h1.grad()
                if drop_off is not None:
                    da2 = a2.grad()
                    fns.append(lambda: ops.dropout_(da2, pl.drop_state, drop_off, pl.drop_p))
                if ssn:
                    fns.append(self.plan_gn_bwd(a2, h1, None, gn2, st2, dh1, film=film, film_off=foff, dfilm=dfilm))
                else:
                    fns.append(self.plan_gn_bwd(a2, h1, None, gn2, st2, dh1))
                    # d emb_out[b, c] = sum over pixels of d(conv1 output)[b, :, :, c]
                    fns.append(lambda: ops.colsum_(dh1, dfilm[:, foff:foff + cout], c=cout, groups=B, out_ld=dfilm.shape[1]))
                h1.g_written = True
                # conv1
                fns += self.plan_conv_bwd(pl, cw1, [a1], dh1, 3)
                # GN1 backward -> dx (+ identity-skip gradient)
                acc0 = x0.grad_acc()
                acc1 = x1.grad_acc() if x1 is not None else False
                dx0 = x0.grad()
                dx1 = x1.grad() if x1 is not None else None
                accmask = (1 if acc0 else 0) | (2 if acc1 else 0)
                dadd = None if has_skip else dout
                fns.append(self.plan_gn_bwd(a1, x0, x1, gn1, st1, dx0, dx1, accmask, dadd=dadd))
                return fns
            pl.bwd_builders.append(build_bwd)
        return out

    def plan_attention(self, pl, ab, x):
        B, H, W, Cc = x.shape
        Tn = H * W
        heads = ab.num_heads
        cwq, cwp = self.convs[id(ab.qkv)], self.convs[id(ab.proj_out)]
        n, qkv, o, out = T(pl, (B, H, W, Cc)), T(pl, (B, H, W, 3 * Cc)), T(pl, (B, H, W, Cc)), T(pl, (B, H, W, Cc))
        st = (pl.alloc((B, 32), th.float32), pl.alloc((B, 32), th.float32))
        lse = pl.alloc((B, heads, Tn), th.float32)
        dsum = pl.alloc((B, heads, Tn), th.float32) if pl.train else None
        self.plan_gn_fwd(pl, x, None, ab.norm, n, st, silu=False)
        self.plan_conv(pl, cwq, [n], qkv, 1)
        qv, ov = qkv.t.view(B, Tn, 3 * Cc), o.t.view(B, Tn, Cc)
        pl.add_fwd(lambda: ops.attn_fwd(qv, heads, out=ov, lse=lse))
        self.plan_conv(pl, cwp, [o], out, 1, resid=x, stats=True)
        if pl.train:
            def build_bwd():
                fns = []
                dout = out.grad()
                fns += self.plan_conv_bwd(pl, cwp, [o], dout, 1)
                dqkv = qkv.grad()
                do = o.grad()
                fns.append(lambda: ops.attn_bwd(qv, ov, do.view(B, Tn, Cc), lse, heads, dqkv=dqkv.view(B, Tn, 3 * Cc),
                                                dsum=dsum))
                qkv.g_written = True
                fns += self.plan_conv_bwd(pl, cwq, [n], dqkv, 1)
                acc = x.grad_acc()
                dx = x.grad()
                fns.append(self.plan_gn_bwd(n, x, None, ab.norm, st, dx, accmask=1 if acc else 0, silu=False, dadd=dout))
                return fns
            pl.bwd_builders.append(build_bwd)
        return out

    def plan_downsample(self, pl, ds, x):
        B, H, W, Cc = x.shape
        cw = self.convs[id(ds.op)]
        out = T(pl, (B, H // 2, W // 2, Cc))
        self.plan_conv(pl, cw, [x], out, 3, stride=2, stats=True)
        if pl.train:
            pl.bwd_builders.append(lambda: self.plan_conv_bwd(pl, cw, [x], out.grad(), 3, stride=2))
        return out

    def plan_upsample(self, pl, up, x):
        B, H, W, Cc = x.shape
        cw = self.convs[id(up.conv)]
        out = T(pl, (B, 2 * H, 2 * W, Cc))
        if not pl.train and UP_ON_LOAD and cw.cout % 128 == 0 and (2 * H) % 32 == 0 and (2 * W) % 8 == 0 and Cc % 64 == 0:
            # inference: the conv expands the low-resolution halo tile itself (make_igemm_desc(up2x=True)): no 4x tensor
            chans = [Cc]
            segs, _ = ops.conv_segments(chans, 3)
            st = out.stats = pl.alloc_stats(B, cw.cout) if FUSED_GN_STATS else None
            d = ops.make_igemm_desc([x.t], segs, cw.fwd, out.t, cw.cout, bias=cw.bias, stats=st, up2x=True)
            pl.add_fwd(lambda: ops.igemm(d))
            return out
        u = T(pl, (B, 2 * H, 2 * W, Cc))
        pl.add_fwd(lambda: ops.upsample2x(x.t, out=u.t))
        self.plan_conv(pl, cw, [u], out, 3, stats=True)
        if pl.train:
            def build_bwd():
                fns = self.plan_conv_bwd(pl, cw, [u], out.grad(), 3)
                du = u.grad()
                acc, dx = x.grad_acc(), x.grad()
                fns.append(lambda: ops.sumpool2x(du, out=dx, accumulate=acc))
                return fns
            pl.bwd_builders.append(build_bwd)
        return out

    # ------------------------------------------------------------------ whole torso
    def build_plan(self, B, train):
        from .unet import ResBlock, AttentionBlock, Upsample, Downsample
        m = self.model
        S = m.image_size
        pl = Plan(self, B, train)
        pl.bwd_builders = []
        if S is None:
            raise _lib.CdaeError("UNetModel needs image_size (create_model passes it)")
        pl.x_in = pl.alloc((B, m.in_channels, S, S), th.float32)
        pl.film_in = pl.alloc((B, self.film_width), th.float32)
        pl.dfilm = pl.alloc((B, self.film_width), th.float32) if train else None
        pl.eps = pl.alloc((B, m.out_channels, S, S), th.float32)
        xin = T(pl, (B, S, S, 64))
        pl.add_fwd(lambda: ops.nchw_to_nhwc_pad(pl.x_in, 64, out=xin.t))
        stem = m.input_blocks[0][0]
        cws = self.convs[id(stem)]
        h = T(pl, (B, S, S, stem.out_channels))
        self.plan_conv(pl, cws, [xin], h, 3, stats=True)
        if train:
            pl.bwd_builders.append(lambda h0=h: self.plan_conv_bwd(pl, cws, [xin], h0.grad(), 3, need_dgrad=False))
        hs = [h]

        def run_block(blk, h, skip=None):
            for mod in blk:
                if isinstance(mod, ResBlock):
                    h = self.plan_resblock(pl, mod, h, skip, pl.film_in, pl.dfilm)
                    skip = None
                elif isinstance(mod, AttentionBlock):
                    h = self.plan_attention(pl, mod, h)
                elif isinstance(mod, Downsample):
                    h = self.plan_downsample(pl, mod, h)
                elif isinstance(mod, Upsample):
                    h = self.plan_upsample(pl, mod, h)
                else:
                    raise _lib.CdaeError(f"unplanned module {type(mod)}")
            return h

        for blk in list(m.input_blocks)[1:]:
            h = run_block(blk, h)
            hs.append(h)
        h = run_block(m.middle_block, h)
        for blk in m.output_blocks:
            h = run_block(blk, h, skip=hs.pop())
        gno, co = m.out[0], m.out[2]
        cwo = self.convs[id(co)]
        sto = (pl.alloc((B, 32), th.float32), pl.alloc((B, 32), th.float32))
        if self.gn_on_load_ok(pl, cwo, h, None):
            self.plan_conv(pl, cwo, [h], pl.eps, 3, out_mode=1, gn_ab=self.plan_gn_constants(pl, h, None, gno, sto))
        else:
            a = T(pl, h.shape)
            self.plan_gn_fwd(pl, h, None, gno, a, sto)
            self.plan_conv(pl, cwo, [a], pl.eps, 3, out_mode=1)
        if train:
            pl.deps_in = pl.alloc((B, m.out_channels, S, S), th.float32)
            dyo = pl.alloc((B, S, S, 64))
            hfin = h

            def build_out_bwd():
                fns = [lambda: ops.nchw_to_nhwc_pad(pl.deps_in, 64, out=dyo)]
                fns += self.plan_conv_bwd(pl, cwo, [a], dyo, 3)
                acc, dx = hfin.grad_acc(), hfin.grad()
                fns.append(self.plan_gn_bwd(a, hfin, None, gno, sto, dx, accmask=1 if acc else 0))
                return fns
            pl.bwd_builders.append(build_out_bwd)
            # backward ops are built in REVERSE layer order so that "first producer writes, later ones accumulate"
            # is decided in true execution order
            built = [b() for b in reversed(pl.bwd_builders)]
            for fns in reversed(built):
                pl.add_bwd(fns)
            pl.add_bwd([pl._zero_bwd_ws])        # layers replay last-to-first: this node opens the backward graph
        return pl

    MAX_PENDING = 2      # training plans of one batch size that may await their backward at the same time

    def plan(self, B, train, fresh=False):
        """The plan of (batch size, train|infer).  A training plan owns the activations its backward reads, so a second
        grad-enabled forward before that backward (two model calls inside one loss, an evaluation forward without
        no_grad) must not reuse it: `fresh=True` returns an instance that is not awaiting a backward, building up to
        MAX_PENDING of them; beyond that the oldest one is recycled and its stale backward raises (generation check)."""
        key = (B, bool(train))
        slots = self.plans.setdefault(key, [])
        if not slots:
            slots.append(self.build_plan(B, train))
        if not (train and fresh):
            return slots[0]
        for pl in slots:
            if not pl.pending:
                return pl
        if len(slots) < self.MAX_PENDING:
            slots.append(self.build_plan(B, train))
            return slots[-1]
        slots.append(slots.pop(0))       # recycle the oldest: its generation moves on, a late backward fails loudly
        return slots[-1]

    # ------------------------------------------------------------------ public entry: the torso as one autograd node
    def film(self, emb):
        """all 22 emb_layers(SiLU(emb)) as one GEMM (ref unet.py:148-154,186) -> [B, sum 2*Cout] fp32 (no gradient:
        standalone layer calls; the model path goes through rep.TrunkRunner)"""
        e = emb.detach().float().contiguous()
        out = th.empty(e.shape[0], self.film_width, device=e.device, dtype=th.float32)
        self.pack()
        return self.film_fwd(e, out, th.empty(e.shape, device=e.device, dtype=bf16))

    def film_fwd(self, emb, out, a_bf):
        """out [B, film_width] fp32 = emb_layers(SiLU(emb)) for all ResBlocks: one tcgen05 GEMM (bf16 operands from the packed
        arena, fp32 accumulate / result); a_bf: bf16 [B, ted] scratch that receives SiLU(emb) (kept for the backward)"""
        fw = self.convs.get("film")
        if fw is None:
            return ops.linear_fwd(emb, self.film_w, self.film_b, out, silu_in=True)
        B, ted = emb.shape
        ops.silu_cast(emb, a_bf, silu=True)
        d = ops.make_igemm_desc([a_bf.view(1, 1, B, ted)], [(0, 0, 0, 0, ted // 64, 0)], fw.fwd, out.view(1, 1, B, self.film_width),
                                self.film_width, bias=self.film_b, out_mode=2)
        ops.igemm(d)
        return out

    def film_bwd(self, emb, a_bf, dfilm, dfilm_bf, demb):
        """gradients of the FiLM projection: film_w.grad / film_b.grad += (tensor-core weight gradient, K = batch), demb [B, ted]
        fp32 = (dfilm Wf) * silu'(emb)"""
        fw = self.convs.get("film")
        if fw is None:
            ops.linear_bwd(emb, self.film_w, dfilm, self.film_w.grad, self.film_b.grad, dx=demb, silu_in=True)
            return ops.silu_bwd_(demb, emb)
        B, ted = emb.shape
        W = self.film_width
        ops.cast_bf16(dfilm, out=dfilm_bf)
        wd = ops.make_wgrad_desc(dfilm_bf.view(1, 1, B, W), a_bf.view(1, 1, B, ted), self.film_w.grad, W, ted, ksize=1,
                                 dbias=self.film_b.grad)
        ops.wgrad(wd)
        d = ops.make_igemm_desc([dfilm_bf.view(1, 1, B, W)], [(0, 0, 0, 0, W // 64, 0)], fw.tr, demb.view(1, 1, B, ted), ted,
                                out_mode=2)
        ops.igemm(d)
        return ops.silu_bwd_(demb, emb)

    @property
    def trunk(self):
        from .rep import TrunkRunner
        if getattr(self, "_trunk", None) is None:
            self._trunk = TrunkRunner(self.model)
        return self._trunk

    def torso_film(self, x, film):
        """eps = torso(x | FiLM vectors of all ResBlocks): ONE autograd node, forward / backward = CUDA-graph replays"""
        if not self.owns(next(self.model.parameters())):
            raise _lib.CdaeError("model parameters were re-homed after the engine was built; access model.engine again")
        train = th.is_grad_enabled() and film.requires_grad
        return _Torso.apply(self, x, film, train)


class _Torso(th.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, film, train):
        B = x.shape[0]
        pl = eng.plan(B, train, fresh=True)
        eng.pack()
        if pl.drop_groups:
            pl.set_dropout(float(eng.model.dropout) if eng.model.training else 0.0)
        pl.x_in.copy_(x)
        pl.film_in.copy_(film)
        pl.forward()
        pl.generation += 1
        pl.pending = bool(train)
        ctx.pl, ctx.generation = pl, pl.generation
        return pl.eps.clone()

    @staticmethod
    def backward(ctx, deps):
        pl = ctx.pl
        if not pl.train:
            raise _lib.CdaeError("backward through a torso forward that ran in inference mode")
        if pl.generation != ctx.generation:
            raise _lib.CdaeError("the saved activations of this torso forward were overwritten by later grad-enabled "
                                 f"forwards of the same batch size (more than {Engine.MAX_PENDING} awaiting backward)")
        pl.pending = False
        pl.deps_in.copy_(deps)
        pl.dfilm.zero_()
        pl.backward()
        return None, None, pl.dfilm.clone(), None


# ---------------------------------------------------------------------- standalone layers (per-layer parity, API users)
def groupnorm_nchw(x, weight, bias, silu=False):
    xs = x.float()
    shp = xs.shape
    x4 = xs.reshape(shp[0], shp[1], -1, 1)
    xn = x4.permute(0, 2, 3, 1).contiguous().to(bf16)
    y, _, _ = ops.gn_fwd(xn, weight.detach().float().contiguous(), bias.detach().float().contiguous(), silu=silu)
    return y.float().permute(0, 3, 1, 2).reshape(shp).to(x.dtype)


def run_layer(mod, x, emb=None):
    """Inference-only execution of ONE torso layer (ResBlock / AttentionBlock / Upsample / Downsample) on NCHW fp32
    input through the same planned kernels the full torso uses."""
    from .unet import ResBlock, AttentionBlock, Upsample, Downsample
    root = getattr(mod, "_cdae_root", None)
    if root is None:
        raise _lib.CdaeError("standalone layer calls need model.engine to have been built (layers are bound to it)")
    eng = root.engine
    eng.pack()
    B, Cc, H, W = x.shape
    pl = Plan(eng, B, False)
    pl.bwd_builders = []
    xin = T(pl, (B, H, W, Cc))
    xin.t.copy_(x.float().permute(0, 2, 3, 1))
    if isinstance(mod, ResBlock):
        film = eng.film(emb)
        out = eng.plan_resblock(pl, mod, xin, None, film.contiguous(), None)
    elif isinstance(mod, AttentionBlock):
        out = eng.plan_attention(pl, mod, xin)
    elif isinstance(mod, Upsample):
        out = eng.plan_upsample(pl, mod, xin)
    elif isinstance(mod, Downsample):
        out = eng.plan_downsample(pl, mod, xin)
    else:
        raise _lib.CdaeError(f"run_layer: unsupported module {type(mod)}")
    pl._run_fwd_eager()
    return out.t.float().permute(0, 3, 1, 2).contiguous()


def _channel_stats(t_bf16):
    """fp32 [B, C, 2] per-(image, channel) sum / sum of squares of an NHWC bf16 tensor: what the producing conv's epilogue
    leaves behind inside the torso (standalone layer calls have no producer)"""
    f = t_bf16.float()
    return th.stack([f.sum(dim=(1, 2)), (f * f).sum(dim=(1, 2))], dim=-1).contiguous()


def run_layer_train(mod, xs, emb=None, dout=None, dropout_p=None):
    """Teacher-forced forward AND backward of ONE torso layer through the same planned kernels (and the same kernel
    variants: statistics epilogue, streaming GroupNorm, two-source concat) the full torso launches.
    xs: NCHW fp32 tensor, or a tuple (h, skip) for the skip-concatenated ResBlocks of the output path; dout: NCHW fp32
    gradient of the layer output.  Returns (out, [dx per input], dfilm [B, film_width] | None), all fp32 NCHW; parameter
    gradients are ACCUMULATED into the engine's gradient arena (read them through `param.grad`)."""
    from .unet import ResBlock, AttentionBlock, Upsample, Downsample
    root = getattr(mod, "_cdae_root", None)
    if root is None:
        raise _lib.CdaeError("standalone layer calls need model.engine to have been built (layers are bound to it)")
    eng = root.engine
    eng.pack()
    xs = list(xs) if isinstance(xs, (tuple, list)) else [xs]
    B, _, H, W = xs[0].shape
    pl = Plan(eng, B, True)
    pl.bwd_builders = []
    tin = []
    for x in xs:
        t = T(pl, (B, H, W, x.shape[1]))
        t.t.copy_(x.float().permute(0, 2, 3, 1))
        if H * W >= 32 and x.shape[1] % 64 == 0 and FUSED_GN_STATS:
            t.stats = _channel_stats(t.t)
        tin.append(t)
    film = dfilm = None
    if isinstance(mod, ResBlock):
        film = eng.film(emb).detach().contiguous()
        dfilm = th.zeros_like(film)
        out = eng.plan_resblock(pl, mod, tin[0], tin[1] if len(tin) > 1 else None, film, dfilm)
    elif isinstance(mod, AttentionBlock):
        out = eng.plan_attention(pl, mod, tin[0])
    elif isinstance(mod, Upsample):
        out = eng.plan_upsample(pl, mod, tin[0])
    elif isinstance(mod, Downsample):
        out = eng.plan_downsample(pl, mod, tin[0])
    else:
        raise _lib.CdaeError(f"run_layer_train: unsupported module {type(mod)}")
    if dropout_p is not None:
        pl.set_dropout(float(dropout_p))
    run_layer_train.last_plan = pl          # tests read the dropout state (to regenerate the masks) from here
    pl._run_fwd_eager()
    res = out.t.float().permute(0, 3, 1, 2).contiguous()
    if dout is None:
        return res, None, None
    built = [b() for b in reversed(pl.bwd_builders)]
    out.grad().copy_(dout.float().permute(0, 2, 3, 1))
    pl._zero_bwd_ws()
    for fns in built:
        for f in fns:
            f()
    dxs = [t.g.float().permute(0, 3, 1, 2).contiguous() for t in tin]
    return res, dxs, dfilm

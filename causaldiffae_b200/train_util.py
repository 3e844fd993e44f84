"""Training driver (ref improved_diffusion/train_util.py): same constructor keywords, methods and attributes.

What changed underneath (B200-first):
  * parameters / gradients / Adam moments / EMA are flat fp32 arenas; AdamW + EMA + sum(g^2) is ONE fused kernel
    (the reference runs ~3 400 per-tensor optimizer ops and 380 `.item()` syncs per step);
  * data parallelism is one NCCL all-reduce (mean) of the contiguous gradient arena over NVLink instead of gloo-DDP
    buckets (ref train_util.py:107-126); parameters are really broadcast from rank 0 at start (ref sync_params is a no-op);
  * nothing on the step path synchronises with the host: losses / grad-norm are accumulated on the device and only
    read when `logger.dumpkvs()` runs (every log_interval steps);
  * rank 0 saves (the reference saves on rank 1 only, so single-process runs never checkpoint, Q8); file names and
    state_dict keys are the reference's, plus the optimizer state the reference forgot.
"""
import math
import os

import numpy as np
import torch as th
import torch.distributed as dist

from . import dist_util, logger, ops
from .resample import LossAwareSampler, UniformSampler

INITIAL_LOG_LOSS_SCALE = 20.0


def adam_hyper(lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, ema_rate=0.9999, grad_scale=1.0):
    """Scalars of one AdamW step (torch.optim.AdamW semantics, ref train_util.py:94) as the fused kernel reads them from
    device memory: {lr, b1, b2, eps, wd, ema_rate, grad_scale}.  The step counter lives on the device and the bias
    corrections are evaluated there (in double, like torch does on the host)."""
    return [lr, beta1, beta2, eps, weight_decay, ema_rate, grad_scale]


# multi-GPU: start the all-reduce of the early gradient range beside the representation path's backward
OVERLAP_ALLREDUCE = os.environ.get("CDAE_OVERLAP_ALLREDUCE", "1") != "0"


class _FlatAdamW:
    """The optimizer object behind `TrainLoop.opt`: torch.optim.AdamW's surface (param_groups / state_dict / step)
    over the flat arena, executed by the fused kernel.  Nothing here synchronises with the host: the step counter, the
    hyper-parameters and the non-finite-gradient guard are device scalars at fixed addresses (graph-replayable)."""

    def __init__(self, engine, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8):
        self.engine = engine
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)]
        self.exp_avg = th.zeros_like(engine.arena)
        self.exp_avg_sq = th.zeros_like(engine.arena)
        dev = engine.device
        self.step_dev = th.zeros(1, device=dev, dtype=th.int64)
        self.hyper = th.zeros(8, device=dev)
        self._hyper_host = None
        self.gsq = th.zeros(1, device=dev)
        self.guard = th.zeros(1, device=dev)

    @property
    def step_count(self):
        return int(self.step_dev.item())

    def set_hyper(self, ema_rate=0.0, grad_scale=1.0):
        """upload the hyper-parameters when they changed (lr annealing, a new world size); otherwise no copy at all"""
        g = self.param_groups[0]
        vals = adam_hyper(g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], ema_rate, grad_scale) + [0.0]
        if vals != self._hyper_host:
            self.hyper.copy_(th.tensor(vals).pin_memory(), non_blocking=True)
            self._hyper_host = vals

    def step(self, ema=None, ema_rate=0.0, grad_scale=1.0, guarded=False, grads=None, lognorm=None):
        """grads: the (all-reduced) gradient buffer, fp32 or bf16 (default: the engine's fp32 gradient arena); guarded:
        skip the whole step on the device when sum(g^2) is not finite (ref train_util.py:277-280); lognorm: device fp32[2]
        += {grad norm, 1} for the logger."""
        self.set_hyper(ema_rate, grad_scale)
        self.launch(ema, guarded, grads, lognorm)

    def launch(self, ema=None, guarded=False, grads=None, lognorm=None):
        e = self.engine
        ops.zero_(self.gsq)
        guard = None
        g = e.grad_arena if grads is None else grads
        if guarded:
            ops.zero_(self.guard)
            ops.sumsq(g, self.guard)
            guard = self.guard
        ops.adam_ema(e.arena, g, self.exp_avg, self.exp_avg_sq, ema, self.hyper, self.step_dev, self.gsq, guard, lognorm)
        e.dirty = True

    def state_dict(self):
        """moments keyed by parameter name with the model's shapes (like the weights and the EMA): the file does not depend on
        the order of the flat arena"""
        e = self.engine
        return dict(step=self.step_count, param_groups=self.param_groups,
                    exp_avg={k: v.clone() for k, v in e.export_state(self.exp_avg).items()},
                    exp_avg_sq={k: v.clone() for k, v in e.export_state(self.exp_avg_sq).items()})

    def load_state_dict(self, sd):
        if th.is_tensor(sd["exp_avg"]):
            raise ValueError("optimizer checkpoint stores flat moment arenas (an older layout of this package): the arena order "
                             "has changed since, re-save it from the version that wrote it as per-parameter tensors")
        self.step_dev.fill_(int(sd["step"]))
        with th.no_grad():
            for key, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                views = self.engine.export_state(flat)
                missing = set(views) - set(sd[key])
                if missing:
                    raise KeyError(f"optimizer checkpoint lacks {key} of {sorted(missing)[:3]} ...")
                for name, v in views.items():
                    v.copy_(sd[key][name])
        self.param_groups = sd["param_groups"]
        self._hyper_host = None


class FusedStep:
    """One training micro-step - q_sample -> conv encoder -> DAG layer -> reparameterisation / keep mask / KL -> embedding
    trunk + FiLM -> UNet torso -> eps-MSE + loss assembly -> the backward of all of it - as explicit launch sequences of
    hand-written kernels over preallocated buffers, replayed as ONE CUDA graph per batch size (ref train_util.py:231-274
    forward_backward + gaussian_diffusion.py:768-859 training_losses + unet.py:525-632 forward + autograd).  No ATen compute
    kernel and no host synchronisation on the path; inputs arrive by plain copies into the static buffers."""

    N_LOG = 24      # [0..19] cdae_step_loss sums, [20..21] gradient-norm sum / count (cdae_adam_ema), rest spare

    @staticmethod
    def supported(loop):
        from .unet import UNetModel
        from . import gaussian_diffusion as gd
        m, d = loop.model, loop.diffusion
        if not isinstance(m, UNetModel) or os.environ.get("CDAE_FUSED_STEP", "1") == "0":
            return False
        if d.model_mean_type not in (gd.ModelMeanType.EPSILON, gd.ModelMeanType.START_X):
            return False
        if d.loss_type not in (gd.LossType.MSE, gd.LossType.RESCALED_MSE):
            return False
        if d.model_var_type in (gd.ModelVarType.LEARNED, gd.ModelVarType.LEARNED_RANGE):
            return False
        return (m.rep_dim is not None) == bool(loop.rep_cond)

    def __init__(self, loop, B):
        from . import gaussian_diffusion as gd
        from .rep import freqs_for
        self.loop, self.B = loop, B
        m, d, eng = loop.model, loop.diffusion, loop.engine
        self.m, self.d, self.eng = m, d, eng
        dev = eng.device
        S, C = m.image_size, m.in_channels
        f = lambda *s: th.zeros(*s, device=dev, dtype=th.float32)   # noqa: E731
        self.x, self.noise = f(B, C, S, S), f(B, C, S, S)
        self.t = th.zeros(B, device=dev, dtype=th.int64)
        self.w = f(B)
        self.y = th.zeros(B, device=dev, dtype=th.int64) if m.num_classes is not None else None
        self.c = None                                              # allocated on first use (label width comes with the data)
        self.klw = f(1)
        self.rng = th.tensor([int(th.initial_seed()) & 0x7fffffffffffffff, 0], device=dev, dtype=th.int64)
        self.rep = m.rep_dim is not None
        self.target_is_noise = d.model_mean_type == gd.ModelMeanType.EPSILON
        self.pl = eng.plan(B, True)
        self.trunk_st = eng.trunk.alloc(B, dev, True)
        if self.rep:
            D = m.rep_dim
            self.enc_st = m.rep_emb.runner.alloc(B, C, S, S, dev, True)
            self.xi, self.keep = f(B, D), (f(B) if m.masking else None)
            self.zp, self.z, self.zpm, self.kld = f(B, D), f(B, D), f(B, D), f(B)
            self.dz, self.dzp, self.dmu, self.dvar, self.du = f(B, D), f(B, D), f(B, D), f(B, D), f(B, D)
            self.dag_ws = f(B, D)
            self.A = m._adjacency(None, dev)
        self.mse, self.loss, self.gscale, self.dkld, self.total = f(B), f(B), f(B), f(B), f(1)
        self.tmap = d._dev_table("timestep_map", dev, lambda: np.asarray(d.timestep_map, dtype=np.int64)) \
            if hasattr(d, "timestep_map") else None
        self.tscale = 1000.0 / d.original_num_steps if getattr(d, "rescale_timesteps", False) and hasattr(d, "timestep_map") else 0.0
        if not hasattr(d, "timestep_map") and d.rescale_timesteps:
            self.tscale = 1000.0 / d.num_timesteps
        self.sqrt_ac = d._f32_table("sqrt_alphas_cumprod", dev)
        self.sqrt_1mac = d._f32_table("sqrt_one_minus_alphas_cumprod", dev)
        freqs_for(m.model_channels, dev)
        self.graph, self.graph_a, self.graph_b, self.runs = None, None, None, 0
        self.device_rng = False

    # ------------------------------------------------------------------ the launch sequence (eager or under capture)
    def _launch(self, part="all"):
        """part "a": everything up to and including the FiLM projection's backward - from there on every gradient in
        [0, engine.early_end) is final; part "b": the rest of the representation path's backward; "all": both.  The split
        lets a multi-GPU run start the all-reduce of the early range (99 % of the arena) beside part b."""
        if part in ("all", "a"):
            self._launch_a()
        if part in ("all", "b"):
            self._launch_b()

    def _launch_a(self):
        m, eng, pl, B = self.m, self.eng, self.pl, self.B
        if self.device_rng:
            ops.randn_(self.noise, self.rng)
            if self.rep:
                ops.randn_(self.xi, self.rng)
                if self.keep is not None:
                    ops.randn_(self.keep, self.rng, bernoulli=True, keep_prob=1 - m.drop_prob)
        from .rep import REP_SIDE_STREAM
        if REP_SIDE_STREAM:
            with ops.on_side():            # the bf16 weight pack (HBM-bound) runs beside the fp32 encoder (small grids)
                eng.pack(force=True)
        else:
            eng.pack(force=True)
        ops.q_sample(self.x, self.noise, self.t, self.sqrt_ac, self.sqrt_1mac, out=pl.x_in)
        z = None
        if self.rep:
            enc = m.rep_emb.runner
            mu, var = enc.forward(self.enc_st, self.x, True)
            if m.causal_modeling:
                pp, _ = m.causal_mask._ptr_tables(want_grads=True)
                ops.dag_fwd(mu, self.A, pp, m.n_vars, m.rep_dim // m.n_vars, m.rep_dim, out=self.zp)
                zp = self.zp
            else:
                zp = mu
            ops.latent_fwd(mu, var, zp, self.xi, self.keep, self.c, self.z, self.zpm, self.kld, m.n_vars,
                           m.causal_modeling, 0.001)
            z = self.z
        if REP_SIDE_STREAM:
            ops.join_side()                # the FiLM projection and the torso read the packed weights
        eng.trunk.forward(self.trunk_st, self.t, self.y, self.c if m.c_dim is not None else None, z, pl.film_in,
                          self.tmap, self.tscale)
        pl._run_fwd_eager()
        target = self.noise if self.target_is_noise else self.x
        ops.mse_loss(pl.eps, target, self.w, want_grad=True, gmul=1.0 / B, mse=self.mse, dpred=pl.deps_in)
        ops.step_loss(self.mse, self.kld if self.rep else None, self.keep if self.rep else None, self.w, self.klw, self.t,
                      self.d.num_timesteps, self.loss, self.gscale, self.dkld, self.total, self.loop.logsums)
        # ---- backward
        ops.zero_(pl.dfilm)
        pl._run_bwd_eager()
        eng.trunk.backward(self.trunk_st, self.y, self.c if m.c_dim is not None else None, z, pl.dfilm,
                           self.dz if self.rep else None, part="film")

    def _launch_b(self):
        m, eng, pl = self.m, self.eng, self.pl
        z = self.z if self.rep else None
        eng.trunk.backward(self.trunk_st, self.y, self.c if m.c_dim is not None else None, z, pl.dfilm,
                           self.dz if self.rep else None, part="rest")
        if self.rep:
            enc = m.rep_emb.runner
            mu, var = self.enc_st.mu, self.enc_st.var
            zp = self.zp if m.causal_modeling else mu
            ops.latent_bwd(mu, var, zp, self.xi, self.keep, self.c, self.dz, self.dkld, None, None, None, self.dzp, self.dmu,
                           self.dvar, m.n_vars, m.causal_modeling, 0.001)
            if m.causal_modeling:
                pp, gp = m.causal_mask._ptr_tables(want_grads=True)
                ops.dag_bwd(mu, self.A, pp, self.dzp, gp, self.dag_ws, m.n_vars, m.rep_dim // m.n_vars, m.rep_dim,
                            du=self.du, du_add=self.dmu)
                dmu = self.du
            else:
                dmu = self.dmu      # dzp IS the gradient w.r.t. mu here: fold it in
                ops.sgemm(dmu, ops._one(dmu.device), self.dzp, 1, dmu.numel(), 1, (0, 0), (dmu.numel(), 1), (dmu.numel(), 1),
                          c_mode=1, splits=1)
            enc.backward(self.enc_st, self.x, dmu, self.dvar)

    def run(self, after_early=None):
        """after_early(event): called once both parts are enqueued; `event` marks the end of part "a" on the current stream
        (the early gradient range is final there).  Without it the step is one graph."""
        from .engine import USE_GRAPHS, _capture
        self.pl.generation += 1            # a pending autograd backward on this plan must fail loudly, not read these buffers
        self.pl.pending = False
        if self.pl.drop_groups:
            self.pl.set_dropout(float(self.m.dropout))
        graphs = USE_GRAPHS and self.runs >= 2
        if after_early is None:
            if graphs:
                if self.graph is None:
                    from . import _lib
                    k0 = _lib.kernel_count()
                    self.graph = _capture(self._launch)
                    self.graph_kernels = _lib.kernel_count() - k0      # kernels one replay of this graph launches (counted)
                self.graph.replay()
            else:
                self._launch()
        else:
            if graphs and self.graph_a is None:
                from . import _lib
                k0 = _lib.kernel_count()
                self.graph_a = _capture(self._launch_a)
                self.graph_b = _capture(self._launch_b)
                self.graph_kernels = _lib.kernel_count() - k0
            ev = th.cuda.Event()
            self.graph_a.replay() if graphs else self._launch_a()
            ev.record()
            self.graph_b.replay() if graphs else self._launch_b()
            after_early(ev)
        self.runs += 1
        self.eng.dirty = False             # the graph packed the current weights; the optimizer marks them dirty again


class TrainLoop:
    def __init__(self, *, model, diffusion, data, batch_size, microbatch, lr, ema_rate, log_interval, save_interval,
                 resume_checkpoint, use_fp16=False, fp16_scale_growth=1e-3, schedule_sampler=None, weight_decay=0.0,
                 lr_anneal_steps=0, rep_cond=False, n_vars=None, causal_modeling=False, flow_based=False,
                 in_channels=3, masking=False):
        if not dist.is_initialized():
            dist_util.setup_dist()
        self.model, self.diffusion, self.data = model, diffusion, data
        self.batch_size = batch_size
        self.microbatch = microbatch if microbatch > 0 else batch_size
        self.lr = lr
        self.ema_rate = [ema_rate] if isinstance(ema_rate, float) else [float(x) for x in ema_rate.split(",")]
        self.log_interval, self.save_interval = log_interval, save_interval
        self.resume_checkpoint = resume_checkpoint
        self.use_fp16, self.fp16_scale_growth = use_fp16, fp16_scale_growth
        self.schedule_sampler = schedule_sampler or UniformSampler(diffusion)
        self.weight_decay, self.lr_anneal_steps = weight_decay, lr_anneal_steps
        self.rep_cond, self.n_vars, self.causal_modeling = rep_cond, n_vars, causal_modeling
        self.flow_based, self.in_channels, self.masking = flow_based, in_channels, masking
        self.step = 0
        self.resume_step = 0
        self.world_size = dist.get_world_size()
        self.global_batch = self.batch_size * self.world_size
        self.lg_loss_scale = INITIAL_LOG_LOSS_SCALE
        self.sync_cuda = th.cuda.is_available()
        self.log_quartiles = True
        self.grad_wire_dtype = th.float32 if os.environ.get("CDAE_GRAD_WIRE", "bf16") == "fp32" else th.bfloat16

        if next(model.parameters()).device.type != "cuda":
            model.to(dist_util.dev())
        self._load_and_sync_parameters()
        self.engine = model.engine
        self.model_params = list(model.parameters())
        self.master_params = [self.engine.arena]        # one flat fp32 master tensor (cf. fp16_util.make_master_params)
        if self.use_fp16:
            self._setup_fp16()
        self.opt = _FlatAdamW(self.engine, lr=self.lr, weight_decay=self.weight_decay)
        if self.resume_step:
            self._load_optimizer_state()
            self.ema_params = [self._load_ema_parameters(rate) for rate in self.ema_rate]
        else:
            self.ema_params = [[self.engine.arena.clone()] for _ in self.ema_rate]
        self.use_ddp = self.world_size > 1
        self.ddp_model = self.model       # data parallelism = flat gradient all-reduce in forward_backward
        self.noise_override = None
        self.logsums = th.zeros(FusedStep.N_LOG, device=self.engine.device)
        self._fused = {}
        self.use_fused = FusedStep.supported(self)
        logger.set_dump_hook("train_step", self._dump_device_sums)
        dist_util.sync_params([self.engine.arena] + [b for b in model.buffers()])

    # ------------------------------------------------------------------ checkpoints
    def _load_and_sync_parameters(self):
        resume_checkpoint = find_resume_checkpoint() or self.resume_checkpoint
        if resume_checkpoint:
            self.resume_step = parse_resume_step_from_filename(resume_checkpoint)
            logger.log(f"loading model from checkpoint: {resume_checkpoint}...")
            self.model.load_state_dict(dist_util.load_state_dict(resume_checkpoint, map_location=dist_util.dev()))

    def _load_ema_parameters(self, rate):
        ema = self.engine.arena.clone()
        main_checkpoint = find_resume_checkpoint() or self.resume_checkpoint
        path = _rank0_decides(lambda: find_ema_checkpoint(main_checkpoint, self.resume_step, rate))
        if path:
            logger.log(f"loading EMA from checkpoint: {path}...")
            sd = dist_util.load_state_dict(path, map_location=dist_util.dev())
            views = self.engine.export_state(ema)
            with th.no_grad():
                for name, v in views.items():
                    v.copy_(sd[name])
        return [ema]

    def _load_optimizer_state(self):
        main_checkpoint = find_resume_checkpoint() or self.resume_checkpoint
        path = os.path.join(os.path.dirname(main_checkpoint), f"opt{self.resume_step:06}.pt")
        if _rank0_decides(lambda: os.path.exists(path)):
            logger.log(f"loading optimizer state from checkpoint: {path}")
            self.opt.load_state_dict(dist_util.load_state_dict(path, map_location=dist_util.dev()))

    def _setup_fp16(self):
        self.model.convert_to_fp16()

    def linear_kl_weight_scheduler(self, step, total_steps, initial, final):
        """ref train_util.py:176-187"""
        if step >= total_steps:
            return final
        if step <= 0:
            return initial
        if total_steps <= 1:
            return final
        t = step / (total_steps - 1)
        return (1.0 - t) * initial + t * final

    # ------------------------------------------------------------------ loop
    def run_loop(self):
        """ref train_util.py:191-219"""
        while not self.lr_anneal_steps or self.step + self.resume_step < self.lr_anneal_steps:
            batch, cond = next(self.data)
            self.run_step(batch, cond)
            if self.step % self.log_interval == 0:
                logger.dumpkvs()
            if self.step % self.save_interval == 0:
                self.save()
                if os.environ.get("DIFFUSION_TRAINING_TEST", "") and self.step > 0:
                    return
            self.step += 1
            self.diffusion.kl_weight = self.linear_kl_weight_scheduler(self.step, 50000, 0.0, 1.0)
        if (self.step - 1) % self.save_interval != 0:
            self.save()

    def run_step(self, batch, cond):
        self.forward_backward(batch, cond)
        if self.use_fp16:
            self.optimize_fp16()
        else:
            self.optimize_normal()
        self.log_step()

    def forward_backward(self, batch, cond):
        """ref train_util.py:231-274"""
        dev = self.engine.device
        ops.zero_(self.engine.grad_arena)
        if self.use_fused and self._fused_ok(batch, cond):
            return self._forward_backward_fused(batch, cond)
        for i in range(0, batch.shape[0], self.microbatch):
            micro = batch[i:i + self.microbatch].to(dev, non_blocking=True)
            micro_cond = {k: v[i:i + self.microbatch].to(dev, non_blocking=True) for k, v in cond.items()}
            t, weights = self.schedule_sampler.sample(micro.shape[0], dev)
            losses = self.diffusion.training_losses(self.ddp_model, micro, t, model_kwargs=micro_cond,
                                                    rep_cond=self.rep_cond, causal_modeling=self.causal_modeling)
            if isinstance(self.schedule_sampler, LossAwareSampler):
                self.schedule_sampler.update_with_local_losses(t, losses["loss"].detach())
            loss = (losses["loss"] * weights).mean()
            self.last_loss = loss.detach()
            log_loss_dict(self.diffusion, t, {k: v * weights for k, v in losses.items()}, self.log_quartiles)
            loss.backward()
        self._grad_scale, self._reduced_grads = 1.0, None
        if self.use_ddp:
            self._grad_scale, self._reduced_grads = self._exchange_gradients()

    def _fused_ok(self, batch, cond):
        extra = set(cond) - {"c", "y"}
        if extra or batch.dim() != 4 or (self.rep_cond and "c" not in cond):
            return False
        m = self.model
        return (("y" in cond) == (m.num_classes is not None)) and (m.c_dim is None or "c" in cond)

    def _forward_backward_fused(self, batch, cond):
        """forward_backward through FusedStep (one CUDA-graph replay per microbatch): the same mathematics as the generic
        path below, without autograd and without a single ATen compute kernel"""
        from . import nn as cnn
        dev = self.engine.device
        if not self.model.training:
            self.model.train()
        for i in range(0, batch.shape[0], self.microbatch):
            micro = batch[i:i + self.microbatch]
            B = micro.shape[0]
            fs = self._fused.get(B)
            if fs is None:
                fs = self._fused[B] = FusedStep(self, B)
            fs.x.copy_(micro, non_blocking=True)
            if "c" in cond:
                cc = cond["c"][i:i + self.microbatch]
                if fs.c is None:
                    fs.c = th.zeros(B, cc.shape[1], device=dev)
                    fs.graph = fs.graph_a = fs.graph_b = None
                fs.c.copy_(cc, non_blocking=True)
            if fs.y is not None:
                fs.y.copy_(cond["y"][i:i + self.microbatch], non_blocking=True)
            if hasattr(self.schedule_sampler, "sample_host"):      # pinned staging -> the static buffers, no kernel
                ti, tw = self.schedule_sampler.sample_host(B)
                fs.t.copy_(th.from_numpy(ti).pin_memory(), non_blocking=True)
                fs.w.copy_(th.from_numpy(tw).pin_memory(), non_blocking=True)
            else:
                t, weights = self.schedule_sampler.sample(B, dev)
                fs.t.copy_(t); fs.w.copy_(weights)
            klw = float(self.diffusion.kl_weight)
            if klw != getattr(fs, "_klw_host", None):
                fs.klw.copy_(th.tensor([klw], dtype=th.float32).pin_memory(), non_blocking=True)
                fs._klw_host = klw
            device_rng = cnn.RNG_MODE == "device"
            if device_rng != fs.device_rng:
                fs.device_rng, fs.graph, fs.graph_a, fs.graph_b = device_rng, None, None, None
            if not device_rng:      # the reference's draws, in its order: noise (device generator), xi then mask (CPU generator)
                if self.noise_override is not None:            # parity runs feed the oracle's noise (training_losses(noise=...))
                    fs.noise.copy_(self.noise_override[i:i + self.microbatch], non_blocking=True)
                else:
                    fs.noise.normal_()
                if fs.rep:
                    fs.xi.copy_(th.randn(fs.xi.shape))
                    if fs.keep is not None:
                        fs.keep.copy_(th.bernoulli(th.zeros(B) + (1 - self.model.drop_prob)))
            overlap = (OVERLAP_ALLREDUCE and self.use_ddp and self.microbatch >= batch.shape[0])
            self._early_work = None
            fs.run(after_early=self._early_exchange if overlap else None)
            if isinstance(self.schedule_sampler, LossAwareSampler):
                self.schedule_sampler.update_with_local_losses(fs.t, fs.loss.detach().clone())
            self.last_loss = fs.total[0]
        self._grad_scale, self._reduced_grads = 1.0, None
        if self.use_ddp:
            self._grad_scale, self._reduced_grads = self._exchange_gradients()

    def _dump_device_sums(self):
        """logger hook: the per-step sums the fused kernels kept on the device -> the reference's log keys (one host read
        per dump instead of ~900 `.item()` syncs per step)"""
        v = self.logsums.tolist()
        self.logsums.zero_()
        out = {}
        if v[3] > 0:
            for j, key in enumerate(("loss", "mse", "kld_rep")):
                if key == "kld_rep" and not self.rep_cond:
                    continue
                out[key] = v[j] / v[3]
                if self.log_quartiles:
                    for q in range(4):
                        if v[16 + q] > 0:
                            out[f"{key}_q{q}"] = v[4 + 4 * j + q] / v[16 + q]
        if v[21] > 0:
            out["grad_norm"] = v[20] / v[21]
        return out

    def _wire_buffer(self):
        if getattr(self, "_wire", None) is None:
            self._wire = th.empty(self.engine.grad_arena.shape, device=self.engine.device, dtype=th.bfloat16)
        return self._wire

    def _early_exchange(self, ev):
        """The gradients in [0, engine.early_end) - everything but the representation path, 99 % of the arena - are final at
        `ev` (end of FusedStep part "a"): their all-reduce starts on a communication stream beside part "b" (the reference
        overlaps bucket by bucket through DDP hooks, train_util.py:107-126)."""
        e = self.engine
        n0 = e.early_end
        if getattr(self, "_comm", None) is None:
            self._comm = th.cuda.Stream()
        self._comm.wait_event(ev)
        with th.cuda.stream(self._comm):
            if self.grad_wire_dtype == th.bfloat16:
                buf = self._wire_buffer()[:n0]
                ops.cast_bf16(e.grad_arena[:n0], out=buf)
            else:
                buf = e.grad_arena[:n0]
            self._early_work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)

    def _exchange_gradients(self):
        """ref train_util.py:107-126 (DDP mean).  bf16 on the wire (default): one cast pass over the arena, ONE NCCL sum
        all-reduce of the 2-byte copy (187 MB instead of 374 MB at cfg2), and the fused optimizer reads that copy
        directly; fp32 (CDAE_GRAD_WIRE=fp32): the all-reduce runs in place on the arena.  When the early range is already
        on its way (_early_exchange) only the tail is exchanged here, then the streams are joined."""
        e = self.engine
        lo = e.early_end if getattr(self, "_early_work", None) is not None else 0
        if self.grad_wire_dtype == th.bfloat16:
            wire = self._wire_buffer()
            if lo < e.n_params:
                ops.cast_bf16(e.grad_arena[lo:], out=wire[lo:])
            scale, red = dp_all_reduce_(wire[lo:]), wire
        else:
            scale, red = dp_all_reduce_(e.grad_arena[lo:]), None
        if lo:
            self._early_work.wait()
            th.cuda.current_stream().wait_stream(self._comm)
            self._early_work = None
        return scale, red

    def optimize_fp16(self):
        """ref train_util.py:276-290.  The guard is kept, on the device: when any gradient is NaN/Inf the fused kernel
        leaves parameters, moments, EMA and the step counter untouched and lg_loss_scale drops by one; otherwise the
        step happens and lg_loss_scale grows by fp16_scale_growth.  bf16 tensor-core compute with fp32 accumulation
        needs no loss scaling, so lg_loss_scale is bookkeeping (logged like the reference) and never multiplies the loss."""
        self.optimize_normal(guarded=True)
        ok = th.isfinite(self.opt.guard[0])
        if not th.is_tensor(self.lg_loss_scale):
            self.lg_loss_scale = th.full((), float(self.lg_loss_scale), device=self.engine.device)
        self.lg_loss_scale = th.where(ok, self.lg_loss_scale + self.fp16_scale_growth, self.lg_loss_scale - 1)

    def optimize_normal(self, guarded=False):
        """ref train_util.py:292-297: grad-norm log, lr anneal, AdamW step, EMA — one fused kernel."""
        self._anneal_lr()
        self.opt.step(ema=self.ema_params[0][0], ema_rate=self.ema_rate[0], grad_scale=getattr(self, "_grad_scale", 1.0),
                      guarded=guarded, grads=getattr(self, "_reduced_grads", None), lognorm=self.logsums[20:22])
        for rate, params in zip(self.ema_rate[1:], self.ema_params[1:]):
            if guarded:     # the extra rates must skip with the step: blend towards the (unchanged) arena is not a no-op
                ok = th.isfinite(self.opt.guard[0])
                params[0].copy_(th.where(ok, params[0] * rate + self.engine.arena * (1 - rate), params[0]))
            else:
                ops.ema_update(params[0], self.engine.arena, rate)
        self._log_grad_norm()

    def _log_grad_norm(self):
        """ref train_util.py:299-303: the fused optimizer kernel keeps the running sum of gradient norms on the device
        (logsums[20:22]); it reaches the logger through the dump hook"""

    def _anneal_lr(self):
        if not self.lr_anneal_steps:
            return
        frac_done = (self.step + self.resume_step) / self.lr_anneal_steps
        for g in self.opt.param_groups:
            g["lr"] = self.lr * (1 - frac_done)

    def log_step(self):
        logger.logkv("step", self.step + self.resume_step)
        logger.logkv("samples", (self.step + self.resume_step + 1) * self.global_batch)
        if self.use_fp16:
            logger.logkv("lg_loss_scale", self.lg_loss_scale)

    # ------------------------------------------------------------------ saving
    def save(self):
        """ref train_util.py:319-345 — same file names/keys; rank 0 writes; optimizer state saved too."""
        step = self.step + self.resume_step
        if dist.get_rank() == 0:
            d = get_blob_logdir()
            os.makedirs(d, exist_ok=True)
            logger.log(f"saving model 0...")
            th.save(self._master_params_to_state_dict(self.master_params), os.path.join(d, f"model{step:06d}.pt"))
            for rate, params in zip(self.ema_rate, self.ema_params):
                logger.log(f"saving model {rate}...")
                sd = self._master_params_to_state_dict(params)
                th.save(sd, os.path.join(d, "ema_checkpoint.pt"))
                th.save(sd, os.path.join(d, f"ema_{rate}_{step:06d}.pt"))
            th.save(self.opt.state_dict(), os.path.join(d, f"opt{step:06d}.pt"))
        if self.world_size > 1:
            dist.barrier()

    def _master_params_to_state_dict(self, master_params):
        state = {k: v.detach().clone().contiguous() for k, v in self.model.state_dict().items()
                 if not k.startswith("_")}
        for name, v in self.engine.export_state(master_params[0]).items():
            state[name] = v.detach().clone().contiguous()
        return state

    def _state_dict_to_master_params(self, state_dict):
        flat = self.engine.arena.clone()
        with th.no_grad():
            for name, v in self.engine.export_state(flat).items():
                v.copy_(state_dict[name])
        return [flat]


def dp_all_reduce_(flat_grads):
    """Data-parallel gradient exchange (ref train_util.py:107-126, DDP mean): ONE sum all-reduce of the contiguous
    gradient arena (NCCL over NVLink on GPUs); the 1/world averaging is returned and folded into the fused optimizer
    kernel's grad_scale instead of costing another pass over the arena."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / world


def parse_resume_step_from_filename(filename):
    """ref train_util.py:366-378"""
    split = filename.split("model")
    if len(split) < 2:
        return 0
    try:
        return int(split[-1].split(".")[0])
    except ValueError:
        return 0


def get_blob_logdir():
    return os.environ.get("DIFFUSION_BLOB_LOGDIR", logger.get_dir())


def find_resume_checkpoint():
    return None


def _rank0_decides(fn):
    """Evaluate a filesystem question on rank 0 and broadcast the answer: `dist_util.load_state_dict` is a collective
    (rank 0 reads, the bytes are broadcast), so every rank must take the same branch even without a shared filesystem."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return fn()
    box = [fn() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def find_ema_checkpoint(main_checkpoint, step, rate):
    """ref train_util.py:386-394: `ema_{rate}_{step}.pt` next to the model checkpoint.  The reference scripts also write
    a rolling `ema_checkpoint.pt` (overwritten by every rate and every save): it is only a fallback, with a warning,
    because it may belong to another step or rate."""
    if main_checkpoint is None:
        return None
    d = os.path.dirname(main_checkpoint)
    path = os.path.join(d, f"ema_{rate}_{step:06d}.pt")
    if os.path.exists(path):
        return path
    path = os.path.join(d, "ema_checkpoint.pt")
    if os.path.exists(path):
        logger.log(f"warning: ema_{rate}_{step:06d}.pt not found; falling back to the rolling {path} "
                   "(may belong to another step or EMA rate)")
        return path
    return None


def log_loss_dict(diffusion, ts, losses, quartiles=True):
    """ref train_util.py:401-407 without host syncs: means and per-timestep-quartile means stay on the device until
    logger.dumpkvs() reads them."""
    q1h = None
    for key, values in losses.items():
        values = values.detach()
        if values.dim() == 0:
            values = values.expand(ts.shape[0])
        logger.logkv_mean(key, values.mean())
        if quartiles:
            if q1h is None:
                q = (4 * ts // diffusion.num_timesteps).clamp_(0, 3)
                q1h = th.nn.functional.one_hot(q, 4).to(values.dtype)       # [B, 4]
                cnt = q1h.sum(0)
            sums = values @ q1h
            logger.logkv_mean_n(f"{key}_q", sums, cnt)

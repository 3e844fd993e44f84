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


def adam_hyper(lr, step, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, ema_rate=0.9999, grad_scale=1.0):
    """Host-side scalars of one AdamW step (torch.optim.AdamW semantics, ref train_util.py:94): the 9 floats the
    fused kernel reads from device memory {lr, b1, b2, eps, wd, lr/bias_corr1, sqrt(bias_corr2), ema_rate, grad_scale}."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    return [lr, beta1, beta2, eps, weight_decay, lr / bc1, math.sqrt(bc2), ema_rate, grad_scale]


class _FlatAdamW:
    """The optimizer object behind `TrainLoop.opt`: torch.optim.AdamW's surface (param_groups / state_dict / step)
    over the flat arena, executed by the fused kernel."""

    def __init__(self, engine, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8):
        self.engine = engine
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)]
        self.exp_avg = th.zeros_like(engine.arena)
        self.exp_avg_sq = th.zeros_like(engine.arena)
        self.step_count = 0
        self.gsq = th.zeros(1, device=engine.device)

    def step(self, ema=None, ema_rate=0.0, grad_scale=1.0):
        g = self.param_groups[0]
        self.step_count += 1
        hyper = th.tensor(adam_hyper(g["lr"], self.step_count, g["betas"][0], g["betas"][1], g["eps"],
                                     g["weight_decay"], ema_rate, grad_scale)).pin_memory().to(self.engine.device,
                                                                                               non_blocking=True)
        self.gsq.zero_()
        ops.adam_ema(self.engine.arena, self.engine.grad_arena, self.exp_avg, self.exp_avg_sq, ema, hyper, self.gsq)
        self.engine.dirty = True

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq,
                    param_groups=self.param_groups)

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups = sd["param_groups"]


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
        path = find_ema_checkpoint(main_checkpoint, self.resume_step, rate)
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
        if os.path.exists(path):
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
        self._grad_scale = dp_all_reduce_(self.engine.grad_arena) if self.use_ddp else 1.0

    def optimize_fp16(self):
        """bf16 tensor-core compute needs no loss scaling: same step as optimize_normal (ref train_util.py:276-290)."""
        self.optimize_normal()
        self.lg_loss_scale += self.fp16_scale_growth

    def optimize_normal(self):
        """ref train_util.py:292-297: grad-norm log, lr anneal, AdamW step, EMA — one fused kernel."""
        self._anneal_lr()
        self.opt.step(ema=self.ema_params[0][0], ema_rate=self.ema_rate[0], grad_scale=getattr(self, "_grad_scale", 1.0))
        for rate, params in zip(self.ema_rate[1:], self.ema_params[1:]):
            ops.ema_update(params[0], self.engine.arena, rate)
        self._log_grad_norm()

    def _log_grad_norm(self):
        logger.logkv_mean("grad_norm", th.sqrt(self.opt.gsq[0]))

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


def find_ema_checkpoint(main_checkpoint, step, rate):
    if main_checkpoint is None:
        return None
    d = os.path.dirname(main_checkpoint)
    for name in (f"ema_{rate}_{step:06d}.pt", "ema_checkpoint.pt"):
        path = os.path.join(d, name)
        if os.path.exists(path):
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

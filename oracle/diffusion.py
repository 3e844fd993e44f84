"""Oracle: Gaussian diffusion forward process, training loss, DDIM sampler, reference train step.

Test infrastructure only (see oracle/__init__.py).  fp32 torch on whatever device the inputs
live on; float64 numpy tables gathered then cast to fp32 exactly like the reference's
`_extract_into_tensor` (gaussian_diffusion.py:938-951).
"""
import numpy as np
import torch

from . import schedules
from . import model as omodel


class Diffusion:
    """SpacedDiffusion restated (ref respace.py:65-124 + gaussian_diffusion.py:104-182).
    EPSILON mean type, FIXED_LARGE variance, plain MSE loss (the only live configuration, Q18)."""

    def __init__(self, steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=False,
                 predict_xstart=False):
        base = schedules.named_beta_schedule(noise_schedule, steps)
        use = schedules.space_timesteps(steps, timestep_respacing if timestep_respacing else [steps])
        betas, self.timestep_map = schedules.respaced_betas(base, use)
        self.tables = schedules.diffusion_tables(betas)
        self.num_timesteps = len(betas)
        self.original_num_steps = steps
        self.rescale_timesteps = rescale_timesteps
        self.predict_xstart = predict_xstart
        self.kl_weight = 0.0

    def extract(self, name, t, ndim):
        """ref gaussian_diffusion.py:938-951: float64 table -> gather at t -> .float() -> broadcast."""
        arr = self.tables[name] if isinstance(name, str) else name
        r = torch.from_numpy(arr).to(t.device)[t].float()
        return r.reshape(-1, *([1] * (ndim - 1)))

    def model_timesteps(self, t):
        """_WrappedModel.__call__ (ref respace.py:119-124)."""
        mt = torch.tensor(self.timestep_map, device=t.device, dtype=t.dtype)[t]
        if self.rescale_timesteps:
            mt = mt.float() * (1000.0 / self.original_num_steps)
        return mt

    def q_sample(self, x0, t, noise):
        """ref gaussian_diffusion.py:201-222."""
        return (self.extract("sqrt_alphas_cumprod", t, x0.ndim) * x0
                + self.extract("sqrt_one_minus_alphas_cumprod", t, x0.ndim) * noise)

    def ddim_step(self, x, t, eps_c, eps_u=None, w=None, eta=0.0, noise=None, clip_denoised=True):
        """p_mean_variance (eps -> clamped x0, ref :277-285,320-325,355-361) + ddim_sample (ref :506-558)."""
        eps = eps_c if w is None else w * eps_c + (1 - w) * eps_u
        nd = x.ndim
        r, s = self.extract("sqrt_recip_alphas_cumprod", t, nd), self.extract("sqrt_recipm1_alphas_cumprod", t, nd)
        x0 = eps if self.predict_xstart else r * x - s * eps
        if clip_denoised:
            x0 = x0.clamp(-1, 1)
        e2 = (r * x - x0) / s
        ab, abp = self.extract("alphas_cumprod", t, nd), self.extract("alphas_cumprod_prev", t, nd)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        if noise is None:
            noise = torch.randn_like(x)
        mean = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * e2
        nz = (t != 0).float().reshape(-1, *([1] * (nd - 1)))
        return mean + nz * sigma * noise, x0

    def ddim_sample_loop(self, eps_fn, x_T, eta=0.0, w=None, clip_denoised=True, noise_fn=None):
        """ddim_sample_loop_progressive (ref :632-680). `eps_fn(x, model_t, uncond)` -> eps."""
        x = x_T
        for i in reversed(range(self.num_timesteps)):
            t = torch.tensor([i] * x.shape[0], device=x.device)
            mt = self.model_timesteps(t)
            with torch.no_grad():
                ec = eps_fn(x, mt, False)
                eu = eps_fn(x, mt, True) if w is not None else None
                nz = noise_fn(i) if noise_fn is not None else torch.zeros_like(x) if eta == 0.0 else None
                x, _ = self.ddim_step(x, t, ec, eu, w, eta, nz, clip_denoised)
        return x


def kl_normal(qm, qv, pm, pv):
    """ref nn.py:440-457."""
    return (0.5 * (torch.log(pv) - torch.log(qv) + qv / pv + (qm - pm).pow(2) / pv - 1)).sum(-1)


def representation_loss(mu, var, z_post, causal_modeling, mask, c):
    """ref gaussian_diffusion.py:718-766.  prior(): mean = (c - 0)/(1 - 0) broadcast over d, variance 1."""
    n = c.shape[1]
    kld = kl_normal(mu, var, torch.zeros_like(mu), torch.ones_like(var))
    if causal_modeling:
        d = mu.shape[1] // n
        zp = z_post.reshape(-1, n, d)
        one = torch.ones_like(zp[:, 0, :])
        for i in range(n):
            kld = kld + kl_normal(zp[:, i, :], one, c[:, i:i + 1].float().expand(-1, d), one)
    if mask is not None:
        kld = torch.sum(kld * mask) / torch.sum(mask)
    return kld


def training_losses(diff, sd, cfg, x0, t, noise, y=None, c=None, rep_cond=True, xi=None, mask_draw=None):
    """GaussianDiffusion.training_losses (ref gaussian_diffusion.py:768-859), MSE/EPSILON branch."""
    x_t = diff.q_sample(x0, t, noise)
    eps, mu, var, z_post, mask = omodel.unet_forward(
        sd, cfg, x_t, diff.model_timesteps(t), y=y, c=c if cfg.c_dim is not None else None,
        x_start=x0 if rep_cond else None, xi=xi, mask_draw=mask_draw)
    terms = {}
    if rep_cond:
        terms["kld_rep"] = representation_loss(mu, var, z_post, cfg.causal_modeling, mask, c)
    target = x0 if diff.predict_xstart else noise
    terms["mse"] = ((target - eps) ** 2).mean(dim=list(range(1, eps.ndim)))
    terms["loss"] = terms["mse"] + diff.kl_weight * terms["kld_rep"] if rep_cond else terms["mse"]
    terms["_aux"] = dict(eps=eps, mu=mu, var=var, z_post=z_post, mask=mask, x_t=x_t)
    return terms


class RefTrainer:
    """One reference optimisation step (ref train_util.py:221-303 + torch.optim.AdamW defaults
    betas (0.9,0.999), eps 1e-8, decoupled weight decay; EMA ref nn.py:503-513; KL-weight schedule :213)."""

    def __init__(self, sd, cfg, diff, lr=1e-4, weight_decay=0.0, ema_rate=0.9999):
        self.sd, self.cfg, self.diff = sd, cfg, diff
        self.names = omodel.trainable_names(cfg)
        for n in self.names:
            sd[n].requires_grad_(True)
        self.opt = torch.optim.AdamW([sd[n] for n in self.names], lr=lr, weight_decay=weight_decay)
        self.ema = {n: sd[n].detach().clone() for n in self.names}
        self.ema_rate, self.step = ema_rate, 0

    def run_step(self, x0, t, noise, weights=None, y=None, c=None, rep_cond=True, xi=None, mask_draw=None):
        for n in self.names:
            self.sd[n].grad = None
        terms = training_losses(self.diff, self.sd, self.cfg, x0, t, noise, y, c, rep_cond, xi, mask_draw)
        w = torch.ones_like(terms["mse"]) if weights is None else weights
        loss = (terms["loss"] * w).mean()
        loss.backward()
        gsq = sum(float((self.sd[n].grad ** 2).sum()) for n in self.names if self.sd[n].grad is not None)
        self.opt.step()
        with torch.no_grad():
            for n in self.names:
                self.ema[n].mul_(self.ema_rate).add_(self.sd[n].detach(), alpha=1 - self.ema_rate)
        self.step += 1
        self.diff.kl_weight = schedules.kl_weight_schedule(self.step)
        return dict(loss=float(loss.detach()), mse=float(terms["mse"].detach().mean()), grad_norm=float(np.sqrt(gsq)),
                    kld=float(terms["kld_rep"].detach().mean()) if "kld_rep" in terms else 0.0)


def counterfactual(diff, sd, cfg, x, noise, xi, do_var, do_value, on="mu", w=None, y=None, eta=0.0):
    """The encode -> intervene -> decode recipe of ref scripts/image_causaldae_test.py:405-436 / 535-594:
    mu,var = encode(x); var := 0.001; [mu slice := value]; causal layer; [z_post slice := value];
    z = reparameterize(z_post, var); x_T = q_sample(x, T'-1, noise); DDIM from x_T conditioned on z."""
    d = cfg.rep_dim // cfg.n_vars
    with torch.no_grad():
        mu, _ = omodel.encoder_encode(sd, cfg, x, training=False)
        var = torch.ones_like(mu) * 0.001
        if on == "mu" and do_var is not None:
            mu[:, do_var * d:(do_var + 1) * d] = do_value
        z_pre = omodel.causal_masking(mu, cfg.A, cfg.n_vars)
        z_post = omodel.nonlinearity_add_back_noise(sd, mu, z_pre, cfg.n_vars)
        if on == "z_post" and do_var is not None:
            z_post[:, do_var * d:(do_var + 1) * d] = do_value
        z = z_post + (var ** 0.5) * xi
        t = torch.full((x.shape[0],), diff.num_timesteps - 1, dtype=torch.long, device=x.device)
        x_T = diff.q_sample(x, t, noise)

        def eps_fn(xc, mt, uncond):
            zz = torch.zeros_like(z) if uncond else z
            return omodel.unet_forward(sd, cfg, xc, mt, y=y, z=zz, training=False)[0]

        return diff.ddim_sample_loop(eps_fn, x_T, eta=eta, w=w), z, x_T

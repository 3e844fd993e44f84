"""Gaussian diffusion process behind the reference API (ref improved_diffusion/gaussian_diffusion.py).

Same class / method names, argument meaning and return conventions as the reference so that
`scripts/image_train.py` and the counterfactual recipe of `scripts/image_causaldae_test.py` run unchanged, but
  * all per-timestep tables live on the device as fp32 (the reference re-uploads the float64 table on every
    `_extract_into_tensor` call, gaussian_diffusion.py:938-951);
  * q_sample, the eps-MSE (+ its gradient) and the DDIM update (+ guidance combine) are single fused kernels;
  * `representation_loss` is vectorised on the device (the reference's `prior()` does B*n_vars host syncs, :718-725).
Float64 numpy tables keep the reference attribute names (`betas`, `alphas_cumprod`, ...) and are bit-identical.
"""
import enum
import math

import numpy as np
import torch as th

from . import ops


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """ref gaussian_diffusion.py:21-45"""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """ref gaussian_diffusion.py:48-65"""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def ddim_coef_table(tables, eta=0.0, clip_denoised=True, predict_xstart=False):
    """Per-step scalars of the fused DDIM kernel, computed exactly like the reference does on the device:
    gather float64 table -> fp32 -> fp32 arithmetic (ref gaussian_diffusion.py:533-556, 355-376).
    Rows: {sqrt_recip_ac, sqrt_recipm1_ac, sqrt(ac_prev), sqrt(1-ac_prev-sigma^2), sigma*[t!=0], clip, xstart, 0}."""
    f = np.float32
    ab, abp = tables["alphas_cumprod"].astype(f), tables["alphas_cumprod_prev"].astype(f)
    T = ab.shape[0]
    sigma = f(eta) * np.sqrt((f(1) - abp) / (f(1) - ab)) * np.sqrt(f(1) - ab / abp)
    out = np.zeros((T, 8), dtype=f)
    out[:, 0] = tables["sqrt_recip_alphas_cumprod"].astype(f)
    out[:, 1] = tables["sqrt_recipm1_alphas_cumprod"].astype(f)
    out[:, 2] = np.sqrt(abp)
    out[:, 3] = np.sqrt(f(1) - abp - sigma ** 2)
    out[:, 4] = sigma * (np.arange(T) != 0).astype(f)
    out[:, 5] = 1.0 if clip_denoised else 0.0
    out[:, 6] = 1.0 if predict_xstart else 0.0
    return out


_TABLE_NAMES = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
                "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                "posterior_mean_coef1", "posterior_mean_coef2")


class GaussianDiffusion:
    """ref gaussian_diffusion.py:104-182 (same constructor keywords, same table attributes)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False,
                 causal_modeling=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.causal_modeling = causal_modeling
        self.kl_weight = 0.0
        self._dev_cache = {}

    # ------------------------------------------------------------------ device-resident tables
    def tables(self):
        return {k: getattr(self, k) for k in _TABLE_NAMES}

    def _dev_table(self, key, device, make):
        k = (key, str(device))
        t = self._dev_cache.get(k)
        if t is None:
            t = th.from_numpy(np.ascontiguousarray(make())).to(device)
            self._dev_cache[k] = t
        return t

    def _f32_table(self, name, device):
        return self._dev_table(name, device, lambda: getattr(self, name).astype(np.float32))

    def _extract(self, name_or_arr, t, broadcast_shape):
        """fp32(round(float64 table))[t] broadcast like ref _extract_into_tensor (:938-951), from the device copy."""
        if isinstance(name_or_arr, str):
            tab = self._f32_table(name_or_arr, t.device)
        else:
            tab = th.from_numpy(name_or_arr.astype(np.float32)).to(t.device)
        res = tab[t]
        while len(res.shape) < len(broadcast_shape):
            res = res[..., None]
        return res.expand(broadcast_shape)

    # ------------------------------------------------------------------ forward process
    def q_mean_variance(self, x_start, t):
        """ref :184-199"""
        mean = self._extract("sqrt_alphas_cumprod", t, x_start.shape) * x_start
        variance = self._extract(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = self._extract("log_one_minus_alphas_cumprod", t, x_start.shape)
        return mean, variance, log_variance

    def q_sample(self, x_start, t, noise=None):
        """ref :201-222 — one fused kernel, tables resident on the device."""
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        x_start = x_start.float().contiguous()
        return ops.q_sample(x_start, noise.float().contiguous(), t.long(),
                            self._f32_table("sqrt_alphas_cumprod", x_start.device),
                            self._f32_table("sqrt_one_minus_alphas_cumprod", x_start.device))

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """ref :224-246"""
        assert x_start.shape == x_t.shape
        mean = (self._extract("posterior_mean_coef1", t, x_t.shape) * x_start
                + self._extract("posterior_mean_coef2", t, x_t.shape) * x_t)
        var = self._extract("posterior_variance", t, x_t.shape)
        logvar = self._extract("posterior_log_variance_clipped", t, x_t.shape)
        return mean, var, logvar

    # ------------------------------------------------------------------ reverse process
    def _model_eps(self, model, x, t, model_kwargs, w):
        """model call(s) of p_mean_variance incl. the guidance pair (ref :277-287). Returns (eps_c, eps_u|None)."""
        ts = self._scale_timesteps(t)
        eps_c = model(x, ts, **model_kwargs)[0]
        if w is None:
            return eps_c, None
        kw = dict(model_kwargs)
        z = kw.get("z")
        rep_dim = z.shape[1] if z is not None else getattr(getattr(model, "model", model), "rep_dim", 64)
        kw["z"] = th.zeros((x.shape[0], rep_dim), device=x.device)   # ref hard-codes width 64 (Q4); we use rep_dim
        eps_u = model(x, ts, **kw)[0]
        return eps_c, eps_u

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, w=None):
        """ref :248-353 (fixed-variance model types; learned sigma is dead code in the reference, Q18)."""
        if model_kwargs is None:
            model_kwargs = {}
        B, C = x.shape[:2]
        assert t.shape == (B,)
        eps_c, eps_u = self._model_eps(model, x, t, model_kwargs, w)
        model_output = eps_c if w is None else w * eps_c + (1 - w) * eps_u
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise NotImplementedError("learned-sigma models are not supported (broken in the reference, SURVEY Q18)")
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            var_np = np.append(self.posterior_variance[1], self.betas[1:])
            model_variance = self._extract(var_np, t, x.shape)
            model_log_variance = self._extract(np.log(var_np), t, x.shape)
        else:
            model_variance = self._extract("posterior_variance", t, x.shape)
            model_log_variance = self._extract("posterior_log_variance_clipped", t, x.shape)

        def process_xstart(v):
            if denoised_fn is not None:
                v = denoised_fn(v)
            return v.clamp(-1, 1) if clip_denoised else v

        if self.model_mean_type == ModelMeanType.PREVIOUS_X:
            pred_xstart = process_xstart(self._predict_xstart_from_xprev(x_t=x, t=t, xprev=model_output))
            model_mean = model_output
        elif self.model_mean_type in (ModelMeanType.START_X, ModelMeanType.EPSILON):
            if self.model_mean_type == ModelMeanType.START_X:
                pred_xstart = process_xstart(model_output)
            else:
                pred_xstart = process_xstart(self._predict_xstart_from_eps(x_t=x, t=t, eps=model_output))
            model_mean, _, _ = self.q_posterior_mean_variance(x_start=pred_xstart, x_t=x, t=t)
        else:
            raise NotImplementedError(self.model_mean_type)
        return {"mean": model_mean, "variance": model_variance, "log_variance": model_log_variance,
                "pred_xstart": pred_xstart}

    def _predict_xstart_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return (self._extract("sqrt_recip_alphas_cumprod", t, x_t.shape) * x_t
                - self._extract("sqrt_recipm1_alphas_cumprod", t, x_t.shape) * eps)

    def _predict_xstart_from_xprev(self, x_t, t, xprev):
        assert x_t.shape == xprev.shape
        return (self._extract(1.0 / self.posterior_mean_coef1, t, x_t.shape) * xprev
                - self._extract(self.posterior_mean_coef2 / self.posterior_mean_coef1, t, x_t.shape) * x_t)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return (self._extract("sqrt_recip_alphas_cumprod", t, x_t.shape) * x_t - pred_xstart) \
            / self._extract("sqrt_recipm1_alphas_cumprod", t, x_t.shape)

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """ref :383-414 (ancestral sampler; not on the benchmarked path — thin torch composition)."""
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        noise = th.randn_like(x)
        nonzero_mask = (t != 0).float().view(-1, *([1] * (len(x.shape) - 1)))
        sample = out["mean"] + nonzero_mask * th.exp(0.5 * out["log_variance"]) * noise
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                      device=None, progress=False):
        final = None
        for sample in self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                     denoised_fn=denoised_fn, model_kwargs=model_kwargs,
                                                     device=device, progress=progress):
            final = sample
        return final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                  model_kwargs=None, device=None, progress=False):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = th.full((shape[0],), i, device=device, dtype=th.long)
            with th.no_grad():
                out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                    model_kwargs=model_kwargs)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ DDIM (hot loop #2)
    def _ddim_table(self, device, eta, clip_denoised):
        key = ("ddim", float(eta), bool(clip_denoised))
        return self._dev_table(key, device, lambda: ddim_coef_table(
            self.tables(), eta, clip_denoised, self.model_mean_type == ModelMeanType.START_X))

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0, w=None):
        """ref :506-558 — model call(s) + ONE fused kernel (guidance combine, x0 clamp, eps re-derivation, update).
        The reference always draws randn_like(x) (Q13); we only draw when eta != 0."""
        if model_kwargs is None:
            model_kwargs = {}
        if denoised_fn is not None or self.model_mean_type == ModelMeanType.PREVIOUS_X:
            return self._ddim_sample_unfused(model, x, t, clip_denoised, denoised_fn, model_kwargs, eta, w)
        eps_c, eps_u = self._model_eps(model, x, t, model_kwargs, w)
        noise = th.randn_like(x) if eta != 0.0 else None
        tab = self._ddim_table(x.device, eta, clip_denoised)
        sample, x0 = ops.ddim_step(x.float().contiguous(), eps_c.float().contiguous(), tab, t.to(th.int32).contiguous(),
                                   eps_u=None if eps_u is None else eps_u.float().contiguous(), w=w, noise=noise,
                                   want_xstart=True)
        return {"sample": sample, "pred_xstart": x0}

    def _ddim_sample_unfused(self, model, x, t, clip_denoised, denoised_fn, model_kwargs, eta, w):
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs, w=w)
        eps = self._predict_eps_from_xstart(x, t, out["pred_xstart"])
        alpha_bar = self._extract("alphas_cumprod", t, x.shape)
        alpha_bar_prev = self._extract("alphas_cumprod_prev", t, x.shape)
        sigma = eta * th.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) * th.sqrt(1 - alpha_bar / alpha_bar_prev)
        noise = th.randn_like(x)
        mean_pred = out["pred_xstart"] * th.sqrt(alpha_bar_prev) + th.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
        nonzero_mask = (t != 0).float().view(-1, *([1] * (len(x.shape) - 1)))
        return {"sample": mean_pred + nonzero_mask * sigma * noise, "pred_xstart": out["pred_xstart"]}

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0):
        """ref :560-596"""
        assert eta == 0.0, "Reverse ODE only for deterministic path"
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        eps = self._predict_eps_from_xstart(x, t, out["pred_xstart"])
        alpha_bar_next = self._extract("alphas_cumprod_next", t, x.shape)
        mean_pred = out["pred_xstart"] * th.sqrt(alpha_bar_next) + th.sqrt(1 - alpha_bar_next) * eps
        return {"sample": mean_pred, "pred_xstart": out["pred_xstart"]}

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                         device=None, progress=False, eta=0.0, w=None):
        """ref :598-630.  With this package's UNet the loop is sampling.DdimRunner (one CUDA graph per step, device-side step
        counter, guidance as one 2B batch); anything else takes the step-by-step path below."""
        from .sampling import DdimRunner
        inner = getattr(model, "model", model) if type(model).__name__ == "_WrappedModel" else model
        if DdimRunner.supported(self, inner, denoised_fn, model_kwargs):
            if device is None:
                device = next(inner.parameters()).device
            x_T = noise if noise is not None else th.randn(*shape, device=device)
            kw = model_kwargs or {}
            rows = DdimRunner.MAX_ROWS // (2 if w is not None else 1)      # torso rows per replay (guidance doubles them)
            outs = []
            with th.no_grad():
                for i in range(0, int(shape[0]), rows):
                    n = min(rows, int(shape[0]) - i)
                    key = (id(inner), n, w is not None, bool(clip_denoised), float(eta))
                    runners = self.__dict__.setdefault("_ddim_runners", {})
                    r = runners.get(key)
                    if r is None or r.eng is not inner.engine:
                        r = runners[key] = DdimRunner(self, inner, n, w is not None, clip_denoised, eta)
                    outs.append(r(x_T[i:i + n], {k: v[i:i + n] for k, v in kw.items()}, w))
            return outs[0] if len(outs) == 1 else th.cat(outs, dim=0)
        final = None
        for sample in self.ddim_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                        denoised_fn=denoised_fn, model_kwargs=model_kwargs,
                                                        device=device, progress=progress, eta=eta, w=w):
            final = sample
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, w=None):
        """ref :632-680"""
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        steps = th.arange(self.num_timesteps, device=device, dtype=th.long)   # one upload, sliced per step
        for i in indices:
            t = steps[i:i + 1].expand(shape[0])
            with th.no_grad():
                out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                       model_kwargs=model_kwargs, eta=eta, w=w)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ losses (hot loop #1)
    def prior(self, scale, label, dim):
        """ref :718-725 vectorised: mean[b,j,:] = (label[b,j]-scale[j][0])/(scale[j][1]-0), var = 1."""
        sc = self._dev_table(("prior_scale", tuple(map(tuple, np.asarray(scale).tolist()))), label.device,
                             lambda: np.asarray(scale, dtype=np.float32))
        mean = ((label.float() - sc[:, 0]) / (sc[:, 1] - 0))[:, :, None].expand(-1, -1, dim)
        return mean, th.ones_like(mean)

    def representation_loss(self, mu, var, z_post, causal_modeling, mask, c):
        """ref :727-766 in closed form: KL(N(mu,var)||N(0,1)) + sum_i KL(N(z_post_i,1)||N(c_i,1))."""
        num_vars = c.shape[1]
        kld = 0.5 * (-th.log(var) + var + mu.pow(2) - 1).sum(-1)
        if causal_modeling:
            d = mu.shape[1] // num_vars
            scale = np.array([[0, 1]] * num_vars)
            pm, _ = self.prior(scale, c, d)
            kld = kld + 0.5 * (z_post.reshape(-1, num_vars, d) - pm).pow(2).sum(dim=(1, 2))
        if mask is not None:
            kld = th.sum(kld * mask) / th.sum(mask)
        return kld

    def training_losses(self, model, x_start, t, model_kwargs=None, noise=None, rep_cond=False, causal_modeling=False):
        """ref :768-859. Returns {"mse", "loss"[, "kld_rep"]}; differentiable w.r.t. the model parameters."""
        from .nn import eps_mse
        if model_kwargs is None:
            model_kwargs = {}
        if noise is None:
            noise = th.randn_like(x_start)
        x_t = self.q_sample(x_start, t, noise=noise)
        terms = {}
        if self.loss_type in (LossType.KL, LossType.RESCALED_KL):
            raise NotImplementedError("VLB losses need learned sigmas (dead code in the reference, SURVEY Q18)")
        if self.loss_type not in (LossType.MSE, LossType.RESCALED_MSE):
            raise NotImplementedError(self.loss_type)
        if rep_cond:
            model_kwargs["x_start"] = x_start                      # the reference mutates the caller's dict too (Q11)
            model_output, mu, var, z_post, mask = model(x_t, self._scale_timesteps(t), **model_kwargs)
            terms["kld_rep"] = self.representation_loss(mu, var, z_post, causal_modeling, mask, model_kwargs["c"])
        else:
            model_output = model(x_t, self._scale_timesteps(t), **model_kwargs)[0]
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise NotImplementedError("learned-sigma models are not supported (broken in the reference, SURVEY Q18)")
        target = {
            ModelMeanType.PREVIOUS_X: lambda: self.q_posterior_mean_variance(x_start=x_start, x_t=x_t, t=t)[0],
            ModelMeanType.START_X: lambda: x_start,
            ModelMeanType.EPSILON: lambda: noise,
        }[self.model_mean_type]()
        assert model_output.shape == target.shape == x_start.shape
        terms["mse"] = eps_mse(model_output, target)
        terms["loss"] = terms["mse"] + self.kl_weight * terms["kld_rep"] if rep_cond else terms["mse"]
        return terms

    # ------------------------------------------------------------------ variational bound (evaluation, ref :682-715, 862-935)
    def _vb_terms_bpd(self, model, x_start, x_t, t, clip_denoised=True, model_kwargs=None):
        """one term of the bound in bits/dim: decoder NLL at t == 0, KL(q(x_{t-1}|x_t,x_0) || p(x_{t-1}|x_t)) elsewhere"""
        from .losses import normal_kl, discretized_gaussian_log_likelihood
        from .nn import mean_flat
        true_mean, _, true_logvar = self.q_posterior_mean_variance(x_start=x_start, x_t=x_t, t=t)
        out = self.p_mean_variance(model, x_t, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
        kl = mean_flat(normal_kl(true_mean, true_logvar, out["mean"], out["log_variance"])) / np.log(2.0)
        nll = -discretized_gaussian_log_likelihood(x_start, means=out["mean"], log_scales=0.5 * out["log_variance"])
        nll = mean_flat(nll) / np.log(2.0)
        return {"output": th.where(t == 0, nll, kl), "pred_xstart": out["pred_xstart"]}

    def _prior_bpd(self, x_start):
        """KL(q(x_T | x_0) || N(0, I)) in bits/dim (ref :862-878)"""
        from .losses import normal_kl
        from .nn import mean_flat
        t = th.full((x_start.shape[0],), self.num_timesteps - 1, device=x_start.device, dtype=th.long)
        qt_mean, _, qt_logvar = self.q_mean_variance(x_start, t)
        return mean_flat(normal_kl(mean1=qt_mean, logvar1=qt_logvar, mean2=0.0, logvar2=0.0)) / np.log(2.0)

    def calc_bpd_loop(self, model, x_start, clip_denoised=True, model_kwargs=None):
        """whole bound, T -> 0 (ref :880-935): {"total_bpd" [N], "prior_bpd" [N], "vb" [N,T], "xstart_mse" [N,T], "mse" [N,T]}"""
        from .nn import mean_flat
        device, B = x_start.device, x_start.shape[0]
        vb, xstart_mse, mse = [], [], []
        for i in range(self.num_timesteps - 1, -1, -1):
            t = th.full((B,), i, device=device, dtype=th.long)
            noise = th.randn_like(x_start)
            x_t = self.q_sample(x_start=x_start, t=t, noise=noise)
            with th.no_grad():
                out = self._vb_terms_bpd(model, x_start=x_start, x_t=x_t, t=t, clip_denoised=clip_denoised,
                                         model_kwargs=model_kwargs)
            vb.append(out["output"])
            xstart_mse.append(mean_flat((out["pred_xstart"] - x_start) ** 2))
            eps = self._predict_eps_from_xstart(x_t, t, out["pred_xstart"])
            mse.append(mean_flat((eps - noise) ** 2))
        vb, xstart_mse, mse = th.stack(vb, dim=1), th.stack(xstart_mse, dim=1), th.stack(mse, dim=1)
        prior_bpd = self._prior_bpd(x_start)
        return {"total_bpd": vb.sum(dim=1) + prior_bpd, "prior_bpd": prior_bpd, "vb": vb, "xstart_mse": xstart_mse, "mse": mse}

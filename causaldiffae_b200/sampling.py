"""Counterfactual generation: encode -> do-intervention -> causal layer -> stochastic encode -> DDIM decode.

The recipe of ref scripts/image_causaldae_test.py:405-436 (intervene on the exogenous code `mu`) and :535-594
(intervene on the endogenous code `z_post`), expressed against this package's public API (the script itself cannot run:
it imports modules that are not in the reference repo).  Batches are independent: ranks shard the batch
and only meet in the final all_gather (ref :438-440)."""
import numpy as np
import torch as th

from . import ops
from .nn import reparameterize


class DdimRunner:
    """The DDIM sampling loop of ref gaussian_diffusion.py:598-680 for this package's UNet, without per-step host work: ONE
    CUDA graph holds a whole step - embedding trunk + FiLM projections -> UNet torso -> fused DDIM update (guidance combine,
    x0 clamp, eps re-derivation, sigma noise) - and is replayed T' times; the step index lives on the device and is
    decremented by the graph itself.  Classifier-free guidance (ref :277-285: two model calls per step) runs the conditional
    and the unconditional branch as ONE batch of 2B through the torso."""

    MAX_ROWS = 512          # torso batch of one replay: bounds the activation arena of the inference plan (~80 MB per row)

    @staticmethod
    def supported(diffusion, model, denoised_fn, model_kwargs):
        from .unet import UNetModel
        from . import gaussian_diffusion as gd
        if not isinstance(model, UNetModel) or denoised_fn is not None:
            return False
        if diffusion.model_mean_type == gd.ModelMeanType.PREVIOUS_X:
            return False
        if diffusion.model_var_type in (gd.ModelVarType.LEARNED, gd.ModelVarType.LEARNED_RANGE):
            return False
        kw = model_kwargs or {}
        if set(kw) - {"z", "y", "c"}:
            return False
        if (model.rep_dim is not None) != ("z" in kw) or (model.num_classes is not None) != ("y" in kw):
            return False
        if model.c_dim is not None and "c" not in kw:
            return False
        p = next(model.parameters())
        return p.is_cuda

    def __init__(self, diffusion, model, B, guided, clip_denoised, eta):
        self.d, self.m, self.B, self.guided, self.eta = diffusion, model, B, guided, float(eta)
        eng = self.eng = model.engine
        dev = self.dev = eng.device
        Bp = self.Bp = 2 * B if guided else B
        S, C = model.image_size, model.in_channels
        self.pl = eng.plan(Bp, False)
        self.st = eng.trunk.alloc(Bp, dev, False)
        self.x = th.zeros(B, C, S, S, device=dev)
        self.z = th.zeros(Bp, model.rep_dim, device=dev) if model.rep_dim is not None else None      # second half stays zero
        self.y = th.zeros(Bp, device=dev, dtype=th.int64) if model.num_classes is not None else None
        self.c = th.zeros(Bp, model.c_dim, device=dev) if model.c_dim is not None else None
        self.noise = th.zeros_like(self.x) if self.eta != 0.0 else None
        self.rng = th.tensor([int(th.initial_seed()) & 0x7fffffffffffffff, 0], device=dev, dtype=th.int64)
        self.step64 = th.zeros(1, device=dev, dtype=th.int64)
        self.step32 = th.zeros(1, device=dev, dtype=th.int32)
        self.coef = diffusion._ddim_table(dev, eta, clip_denoised)
        self.tmap = diffusion._dev_table("timestep_map", dev, lambda: np.asarray(diffusion.timestep_map, dtype=np.int64)) \
            if hasattr(diffusion, "timestep_map") else None
        base = getattr(diffusion, "original_num_steps", diffusion.num_timesteps)
        self.tscale = 1000.0 / base if diffusion.rescale_timesteps else 0.0
        self.graphs = {}

    def _launch(self, w):
        B, pl = self.B, self.pl
        pl.x_in[:B].copy_(self.x)
        if self.guided:
            pl.x_in[B:].copy_(self.x)
        self.eng.trunk.forward(self.st, self.step64, self.y, self.c, self.z, pl.film_in, self.tmap, self.tscale)
        pl._run_fwd_eager()
        if self.noise is not None:
            ops.randn_(self.noise, self.rng)
        ops.ddim_step(self.x, pl.eps[:B], self.coef, self.step32, eps_u=pl.eps[B:] if self.guided else None,
                      w=w if self.guided else None, noise=self.noise, out=self.x)
        ops.step_tick(self.step64, self.step32, -1)

    def __call__(self, x_T, kw, w):
        from .engine import USE_GRAPHS, _capture
        B, T = self.B, self.d.num_timesteps
        self.eng.pack()
        self.x.copy_(x_T)
        if self.z is not None:
            self.z[:B].copy_(kw["z"])
        if self.y is not None:
            self.y[:B].copy_(kw["y"]); self.y[B:].copy_(kw["y"][:self.Bp - B])
        if self.c is not None:
            self.c[:B].copy_(kw["c"]); self.c[B:].copy_(kw["c"][:self.Bp - B])
        self.step64.fill_(T - 1); self.step32.fill_(T - 1)
        key = float(w) if self.guided else None
        done = 0
        if USE_GRAPHS and key not in self.graphs:
            self._launch(key)                            # the first step runs eagerly, then the same launches are captured
            done = 1
            self.graphs[key] = _capture(lambda: self._launch(key))
        for _ in range(T - done):
            if USE_GRAPHS:
                self.graphs[key].replay()
            else:
                self._launch(key)
        return self.x.clone()


@th.no_grad()
def encode(model, x, do_var=None, do_value=0.0, on="mu", A=None):
    """steps 1-4 of the recipe: returns (z, mu, z_post) with var := 0.001 as the scripts do (ref :405-413)."""
    mu, var = model.rep_emb.encode(x)
    var = th.ones_like(var) * 0.001
    d = model.rep_dim // model.n_vars
    if do_var is not None and on == "mu":
        mu[:, do_var * d:(do_var + 1) * d] = do_value
    if model.causal_modeling:
        At = th.as_tensor(A if A is not None else model.A, dtype=th.float32, device=mu.device)
        z_post = model.causal_mask(mu, At)       # causal_masking + nonlinearity_add_back_noise (one fused kernel)
    else:
        z_post = mu
    if do_var is not None and on == "z_post":
        z_post[:, do_var * d:(do_var + 1) * d] = do_value
    z = reparameterize(z_post, var)
    return z, mu, z_post


@th.no_grad()
def counterfactual(model, diffusion, x, do_var=None, do_value=0.0, on="mu", w=None, y=None, noise=None, eta=0.0, A=None,
                   clip_denoised=True):
    """x: [B,C,H,W] in [0,1] on the model's device -> counterfactual images [B,C,H,W] (ref :415-436)."""
    z, _, _ = encode(model, x, do_var, do_value, on, A)
    B = x.shape[0]
    t = th.full((B,), diffusion.num_timesteps - 1, device=x.device, dtype=th.long)
    x_T = diffusion.q_sample(x, t, noise=noise if noise is not None else th.randn_like(x))
    cond = {"z": z}
    if y is not None:
        cond["y"] = y
    return diffusion.ddim_sample_loop(model, tuple(x.shape), noise=x_T, clip_denoised=clip_denoised, model_kwargs=cond,
                                      eta=eta, w=w)


@th.no_grad()
def intervention_sweep(model, diffusion, x, do_var, values, on="mu", w=None, y=None, noise=None):
    """'traversal' mode of the reference (ref :481-530): the same x_T decoded under a sweep of do() values."""
    noise = th.randn_like(x) if noise is None else noise
    return th.stack([counterfactual(model, diffusion, x, do_var, float(v), on, w, y, noise) for v in values], dim=1)

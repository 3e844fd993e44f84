"""Counterfactual generation: encode -> do-intervention -> causal layer -> stochastic encode -> DDIM decode.

The recipe of ref scripts/image_causaldae_test.py:405-436 (intervene on the exogenous code `mu`) and :535-594
(intervene on the endogenous code `z_post`), expressed against this package's public API (the script itself cannot run:
it imports modules that are not in the reference repo).  Batches are independent: ranks shard the batch
and only meet in the final all_gather (ref :438-440)."""
import torch as th

from .nn import reparameterize


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

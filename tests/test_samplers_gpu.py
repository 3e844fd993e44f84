"""Samplers on the GPU (SURVEY 8f N4 + the DDIM hot loop):
  * p_mean_variance / p_sample / p_sample_loop (ancestral), ddim_reverse_sample (+ the inversion chain), posterior and
    bits-per-dim pieces against the REAL reference's golden values (tests/golden/samplers_v1.npz) with the stand-in model,
    now with every tensor on the device (the device-resident fp32 tables, the fused q_sample kernel);
  * sampling.DdimRunner - one CUDA graph per DDIM step, device-side step counter, classifier-free guidance as one 2B batch -
    against the step-by-step ddim_sample path of the same model (same kernels, per-step Python loop, two B-sized calls)."""
import os

import numpy as np
import pytest
import torch

from tests.golden import sampler_cases as sc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = dict(rtol=2e-5, atol=2e-5)


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", list(sc.DIFFUSIONS))
def test_samplers_on_device_match_reference_golden(name):
    from causaldiffae_b200 import script_util as su
    gold = np.load(os.path.join(ROOT, "tests", "golden", "samplers_v1.npz"))
    d = su.create_gaussian_diffusion(**sc.DIFFUSIONS[name])
    x, t = sc.inputs(d.num_timesteps)
    x, t = x.cuda(), t.cuda()
    pm = d.p_mean_variance(sc.stub_model, x, t, clip_denoised=True)
    for k in ("mean", "variance", "log_variance", "pred_xstart"):
        np.testing.assert_allclose(_np(pm[k]), gold[f"{name}/pmv/{k}"], **TOL, err_msg=k)
    rv = d.ddim_reverse_sample(sc.stub_model, x, t)
    np.testing.assert_allclose(_np(rv["sample"]), gold[f"{name}/ddim_reverse/sample"], **TOL)
    m, v, lv = d.q_posterior_mean_variance(x * 0.5, x, t)
    np.testing.assert_allclose(_np(m), gold[f"{name}/q_post/mean"], **TOL)
    np.testing.assert_allclose(_np(lv), gold[f"{name}/q_post/logvar"], **TOL)
    # p_sample: mean / pred_xstart are noise-free; the sample adds sigma * N(0,1) from the DEVICE generator, so compare its
    # deterministic part and the noise scale
    ps = d.p_sample(sc.stub_model, x, t)
    np.testing.assert_allclose(_np(ps["pred_xstart"]), gold[f"{name}/p_sample/pred_xstart"], **TOL)
    resid = (ps["sample"] - pm["mean"]) / torch.exp(0.5 * pm["log_variance"])
    nz = (t != 0)
    assert float(resid[~nz].abs().max()) == 0.0 if bool((~nz).any()) else True
    if bool(nz.any()):
        assert 0.8 < float(resid[nz].std()) < 1.2
    if d.num_timesteps <= 20:
        xs = x
        for i in range(d.num_timesteps):      # DDIM inversion chain (deterministic): the reference's values end to end
            xs = d.ddim_reverse_sample(sc.stub_model, xs, torch.full((x.shape[0],), i, dtype=torch.long, device="cuda"))["sample"]
        np.testing.assert_allclose(_np(xs), gold[f"{name}/ddim_reverse_chain"], rtol=5e-5, atol=5e-5)
        out = d.p_sample_loop(sc.stub_model, tuple(x.shape), noise=x, device="cuda")
        assert out.shape == x.shape and bool(torch.isfinite(out).all())
        # DDIM (eta 0) through the fused update kernel, stub model: deterministic -> equals the formula path
        a = d.ddim_sample_loop(sc.stub_model, tuple(x.shape), noise=x, device="cuda")
        b = x
        for i in reversed(range(d.num_timesteps)):
            tt = torch.full((x.shape[0],), i, dtype=torch.long, device="cuda")
            b = d._ddim_sample_unfused(sc.stub_model, b, tt, True, None, {}, 0.0, None)["sample"]
        np.testing.assert_allclose(_np(a), _np(b), rtol=1e-5, atol=1e-5)


CFG = dict(image_size=32, num_channels=64, num_res_blocks=1, rep_cond=True, n_vars=4, causal_modeling=True, in_channels=3,
           learn_sigma=False, rescale_learned_sigmas=False, diffusion_steps=1000)


@pytest.mark.parametrize("class_cond,rescale,w,eta", [(False, False, None, 0.0), (True, True, 1.5, 0.0), (False, False, 0.0, 0.0),
                                                      (False, False, 2.0, 0.5)])
def test_ddim_runner_matches_step_by_step_path(class_cond, rescale, w, eta, monkeypatch):
    from causaldiffae_b200 import script_util as su, sampling
    from oracle import model as om
    full = {**su.model_and_diffusion_defaults(), **CFG, "class_cond": class_cond, "rescale_timesteps": rescale,
            "timestep_respacing": "ddim5"}
    model, diff = su.create_model_and_diffusion(**full)
    cfg = om.config_from_flags(**full)
    model.load_state_dict(om.seeded_state_dict(cfg, seed=0), strict=True)
    model.cuda().eval()
    g = torch.Generator().manual_seed(3)
    B = 6
    x_T = torch.randn(B, 3, 32, 32, generator=g).cuda()
    kw = dict(z=torch.randn(B, 512, generator=g).cuda())
    if class_cond:
        kw["y"] = torch.randint(0, 10, (B,), generator=g).cuda()
    with torch.no_grad():
        if eta == 0.0:
            ref = None
            for out in diff.ddim_sample_loop_progressive(model, tuple(x_T.shape), noise=x_T, model_kwargs=dict(kw), eta=eta, w=w):
                ref = out["sample"]
        fast = diff.ddim_sample_loop(model, tuple(x_T.shape), noise=x_T, model_kwargs=dict(kw), eta=eta, w=w)
        again = diff.ddim_sample_loop(model, tuple(x_T.shape), noise=x_T, model_kwargs=dict(kw), eta=eta, w=w)   # graph replay only
        monkeypatch.setattr(sampling.DdimRunner, "MAX_ROWS", 4 if w is None else 8)      # chunks of 4 rows: 6 = 4 + 2
        chunked = diff.ddim_sample_loop(model, tuple(x_T.shape), noise=x_T, model_kwargs=dict(kw), eta=eta, w=w)
    assert fast.shape == x_T.shape and bool(torch.isfinite(fast).all())
    rel = lambda a, b: float((a - b).norm() / b.norm())      # noqa: E731
    if eta == 0.0:
        # same kernels, different batch geometry (2B / chunks) and fp32-atomic order: bf16 noise floor over 5 (x2) forwards
        assert rel(fast, ref) < 3e-2, rel(fast, ref)
        assert rel(again, ref) < 3e-2 and rel(chunked, ref) < 3e-2, (rel(again, ref), rel(chunked, ref))
        assert rel(ref, x_T) > 0.05
    else:
        assert float(fast.std()) > 0.1 and not torch.equal(fast, again)       # sigma noise is drawn on the device, per call

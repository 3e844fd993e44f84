"""Generate tests/golden/samplers_v1.npz: outputs of the REAL reference GaussianDiffusion / SpacedDiffusion sampler
methods that sit next to the DDIM hot loop - p_mean_variance, p_sample, p_sample_loop,
ddim_reverse_sample, q_posterior_mean_variance, q_mean_variance - driven by a closed-form stand-in model so that no
UNet is involved.  Run in the build container only:  python tests/golden/make_samplers_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402
from tests.golden import sampler_cases as sc  # noqa: E402


def main():
    ns = refshim.load()
    out = {}
    for name, kw in sc.DIFFUSIONS.items():
        d = ns.su.create_gaussian_diffusion(**kw)
        x, t = sc.inputs(d.num_timesteps)
        pm = d.p_mean_variance(sc.stub_model, x, t, clip_denoised=True)
        for k, v in pm.items():
            out[f"{name}/pmv/{k}"] = v.numpy()
        pm = d.p_mean_variance(sc.stub_model, x, t, clip_denoised=False, denoised_fn=sc.denoised_fn)
        out[f"{name}/pmv_fn/mean"] = pm["mean"].numpy()
        torch.manual_seed(11)
        ps = d.p_sample(sc.stub_model, x, t)
        out[f"{name}/p_sample/sample"] = ps["sample"].numpy()
        out[f"{name}/p_sample/pred_xstart"] = ps["pred_xstart"].numpy()
        rv = d.ddim_reverse_sample(sc.stub_model, x, t)
        out[f"{name}/ddim_reverse/sample"] = rv["sample"].numpy()
        m, v, lv = d.q_posterior_mean_variance(x * 0.5, x, t)
        out[f"{name}/q_post/mean"], out[f"{name}/q_post/var"], out[f"{name}/q_post/logvar"] = m.numpy(), v.numpy(), lv.numpy()
        m, v, lv = d.q_mean_variance(x, t)
        out[f"{name}/q_mv/mean"], out[f"{name}/q_mv/var"], out[f"{name}/q_mv/logvar"] = m.numpy(), v.numpy(), lv.numpy()
        vbt = d._vb_terms_bpd(sc.stub_model, x * 0.5, x, t)
        out[f"{name}/vb/output"], out[f"{name}/vb/pred_xstart"] = vbt["output"].numpy(), vbt["pred_xstart"].numpy()
        out[f"{name}/prior_bpd"] = d._prior_bpd(x * 0.5).numpy()
        if d.num_timesteps <= 20:
            torch.manual_seed(13)
            bpd = d.calc_bpd_loop(sc.stub_model, x.clamp(-1, 1))
            for k, v in bpd.items():
                out[f"{name}/bpd/{k}"] = v.numpy()
            torch.manual_seed(12)
            out[f"{name}/p_sample_loop"] = d.p_sample_loop(sc.stub_model, tuple(x.shape), noise=x, device="cpu").numpy()
            xs = x
            for i in range(d.num_timesteps):        # DDIM inversion (encode) over the whole respaced chain
                xs = d.ddim_reverse_sample(sc.stub_model, xs, torch.full((x.shape[0],), i, dtype=torch.long))["sample"]
            out[f"{name}/ddim_reverse_chain"] = xs.numpy()
    # LossSecondMomentResampler (resample.py:122-156): the reference needs `np.int`, removed in numpy 1.24
    if not hasattr(np, "int"):
        np.int = int
    d = ns.su.create_gaussian_diffusion(steps=sc.RESAMPLER["T"])
    smp = ns.resample.LossSecondMomentResampler(d, history_per_term=sc.RESAMPLER["history"], uniform_prob=sc.RESAMPLER["uniform_prob"])
    for r, (ts, losses) in enumerate(sc.resampler_batches()):
        smp.update_with_all_losses(ts, losses)
        if r in sc.RESAMPLER["check_rounds"]:
            out[f"resampler/weights/{r}"] = np.asarray(smp.weights(), dtype=np.float64)
            np.random.seed(100 + r)
            t, w = smp.sample(16, torch.device("cpu"))
            out[f"resampler/t/{r}"], out[f"resampler/w/{r}"] = t.numpy(), w.numpy()
    out["resampler/history"] = smp._loss_history.copy()
    path = os.path.join(HERE, "samplers_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", len(out), "arrays")


if __name__ == "__main__":
    main()

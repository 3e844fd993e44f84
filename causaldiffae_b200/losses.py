"""Likelihood helpers of the variational bound (ref improved_diffusion/losses.py:12-77), used by
GaussianDiffusion._vb_terms_bpd / _prior_bpd / calc_bpd_loop (the `scripts/image_nll.py` evaluation next to the hot path).
Plain fp32 torch on whatever device the inputs live on."""
import math

import torch as th


def normal_kl(mean1, logvar1, mean2, logvar2):
    """KL(N(mean1, exp(logvar1)) || N(mean2, exp(logvar2))), elementwise with broadcasting; scalars are allowed for any
    argument as long as one of them is a tensor (ref :12-37)"""
    anchor = next((v for v in (mean1, logvar1, mean2, logvar2) if isinstance(v, th.Tensor)), None)
    assert anchor is not None, "at least one argument must be a Tensor"
    lv1, lv2 = (v if isinstance(v, th.Tensor) else th.tensor(v).to(anchor) for v in (logvar1, logvar2))
    return 0.5 * (-1.0 + lv2 - lv1 + th.exp(lv1 - lv2) + ((mean1 - mean2) ** 2) * th.exp(-lv2))


def approx_standard_normal_cdf(x):
    """tanh approximation of the standard normal CDF (ref :40-45)"""
    return 0.5 * (1.0 + th.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * th.pow(x, 3))))


def discretized_gaussian_log_likelihood(x, *, means, log_scales):
    """log-probability (nats) of images x in [-1, 1] quantised to 256 levels under N(means, exp(log_scales)^2):
    CDF mass of the 2/255-wide bin around x, open-ended at both extremes (ref :48-77)"""
    assert x.shape == means.shape == log_scales.shape
    centred = x - means
    inv_std = th.exp(-log_scales)
    cdf_hi = approx_standard_normal_cdf(inv_std * (centred + 1.0 / 255.0))
    cdf_lo = approx_standard_normal_cdf(inv_std * (centred - 1.0 / 255.0))
    log_hi = th.log(cdf_hi.clamp(min=1e-12))
    log_upper_tail = th.log((1.0 - cdf_lo).clamp(min=1e-12))
    log_bin = th.log((cdf_hi - cdf_lo).clamp(min=1e-12))
    out = th.where(x < -0.999, log_hi, th.where(x > 0.999, log_upper_tail, log_bin))
    assert out.shape == x.shape
    return out

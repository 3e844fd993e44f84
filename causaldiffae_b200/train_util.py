"""Training driver (ref improved_diffusion/train_util.py) — filled in below; adam_hyper is shared with the tests."""
import math


def adam_hyper(lr, step, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, ema_rate=0.9999, grad_scale=1.0):
    """Host-side scalars of one AdamW step (torch.optim.AdamW semantics, ref train_util.py:94): the 9 floats the
    fused kernel reads from device memory {lr, b1, b2, eps, wd, lr/bias_corr1, sqrt(bias_corr2), ema_rate, grad_scale}."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    return [lr, beta1, beta2, eps, weight_decay, lr / bc1, math.sqrt(bc2), ema_rate, grad_scale]

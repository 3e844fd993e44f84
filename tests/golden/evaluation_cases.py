"""Seeded inputs shared by make_evaluation_golden.py (reference side) and tests/test_evaluation_*.py (this repo)."""
import numpy as np
import torch

DCI_CASES = {
    "codes16_factors4": dict(seed=0, codes=16, factors=4, n_train=400, n_test=150),
    "codes8_factors2": dict(seed=3, codes=8, factors=2, n_train=300, n_test=100),
}


def dci_inputs(case):
    """codes that depend on the factors through a sparse mixing matrix + noise: [codes, N], [factors, N]"""
    r = np.random.RandomState(case["seed"])
    F, Cn = case["factors"], case["codes"]
    mix = r.randn(Cn, F) * (r.rand(Cn, F) < 0.4)
    def draw(n):
        y = r.rand(F, n)
        x = mix @ y + 0.05 * r.randn(Cn, n)
        return x, y
    xtr, ytr = draw(case["n_train"])
    xte, yte = draw(case["n_test"])
    return xtr, ytr, xte, yte


def fixed_importance():
    r = np.random.RandomState(7)
    m = r.rand(12, 4) ** 3
    m[5] = 0.0
    return m


CLF_CASES = {
    "pendulum96": dict(seed=1, in_channels=4, num_vars=4, size=96, batch=5,
                       grad_probe=["encoder.0.0.weight", "encoder.3.1.weight", "encoder.5.1.bias", "fc.weight", "fc.bias"]),
    "mnist28": dict(seed=2, in_channels=1, num_vars=2, size=28, batch=6,
                    grad_probe=["encoder.0.0.weight", "encoder.2.1.weight", "fc.weight"]),
}


def clf_state_dict(template, seed):
    """deterministic weights for every key of the module's state_dict (shapes from the template)"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in template.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros_like(v)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            sd[k] = torch.randn(v.shape, generator=g) * fan_in ** -0.5
        elif k.endswith(".1.weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
    return sd


def clf_inputs(case):
    g = torch.Generator().manual_seed(case["seed"] + 100)
    x = torch.rand(case["batch"], case["in_channels"], case["size"], case["size"], generator=g)
    target = torch.rand(case["batch"], generator=g)
    return x, target

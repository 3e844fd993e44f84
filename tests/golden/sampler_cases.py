"""Seeded inputs and the closed-form stand-in model shared by make_samplers_golden.py (reference side) and
tests/test_samplers_cpu.py (this repo's side)."""
import torch

DIFFUSIONS = {
    "lin1000": dict(steps=1000, noise_schedule="linear"),
    "lin1000_r10": dict(steps=1000, noise_schedule="linear", timestep_respacing="10"),
    "cos100_ddim20_small": dict(steps=100, noise_schedule="cosine", timestep_respacing="ddim20", sigma_small=True),
    "lin1000_r10_xstart": dict(steps=1000, noise_schedule="linear", timestep_respacing="10", predict_xstart=True),
    "lin50_rescaled": dict(steps=50, noise_schedule="linear", rescale_timesteps=True),
}


def inputs(T, B=6, shape=(3, 8, 8)):
    g = torch.Generator().manual_seed(T + 1)
    x = torch.randn(B, *shape, generator=g)
    t = torch.tensor([0, 1, T // 3, T // 2, T - 2, T - 1], dtype=torch.long)[:B]
    return x, t


def stub_model(x, t, **kwargs):
    """eps(x, t): smooth closed form; returns the reference model's 5-tuple"""
    e = 0.3 * x + 0.1 * torch.cos(t.float() / 100.0).view(-1, 1, 1, 1) + 0.05 * torch.sin(3.0 * x)
    return e, None, None, None, None


def denoised_fn(v):
    return torch.tanh(v)


RESAMPLER = dict(T=20, history=3, uniform_prob=0.01, check_rounds=(0, 5, 11, 12, 29))


def resampler_batches(rounds=30, B=8):
    """seeded (timesteps, losses) batches; early rounds leave some timesteps unseen (not warmed up), duplicates occur"""
    g = torch.Generator().manual_seed(77)
    for r in range(rounds):
        hi = 12 if r < 4 else 20
        ts = torch.randint(0, hi, (B,), generator=g).tolist()
        if r >= 4:
            ts[:4] = [(4 * r + k) % 20 for k in range(4)]      # sweep so that every timestep fills its history
        losses = (torch.rand(B, generator=g) * (1 + 0.1 * r)).tolist()
        yield ts, losses

"""Timestep samplers (ref improved_diffusion/resample.py).  Only the uniform sampler is live in the reference
(LossSecondMomentResampler crashes on numpy >= 1.24).  Sampling stays on the host numpy global RNG so
that, given the same np.random.seed, the drawn timesteps are bit-identical to the reference's."""
from abc import ABC, abstractmethod

import numpy as np
import torch as th


def create_named_schedule_sampler(name, diffusion):
    """ref resample.py:9-21"""
    if name == "uniform":
        return UniformSampler(diffusion)
    if name == "loss-second-moment":
        raise NotImplementedError("loss-second-moment resampling is broken in the reference (np.int, SURVEY Q9)")
    raise NotImplementedError(f"unknown schedule sampler: {name}")


class ScheduleSampler(ABC):
    @abstractmethod
    def weights(self):
        """numpy array of positive per-timestep weights"""

    def sample_host(self, batch_size):
        """ref resample.py:44-60 up to the device copy: returns (int64 indices, float32 weights) numpy arrays."""
        w = self.weights()
        p = w / np.sum(w)
        idx = np.random.choice(len(p), size=(batch_size,), p=p)
        return idx.astype(np.int64), (1 / (len(p) * p[idx])).astype(np.float32)

    def sample(self, batch_size, device):
        idx, wts = self.sample_host(batch_size)
        ti, tw = th.from_numpy(idx).long(), th.from_numpy(wts).float()
        if th.device(device).type == "cuda":    # pinned staging + async copy: no stream synchronisation per step
            return ti.pin_memory().to(device, non_blocking=True), tw.pin_memory().to(device, non_blocking=True)
        return ti.to(device), tw.to(device)


class UniformSampler(ScheduleSampler):
    def __init__(self, diffusion):
        self.diffusion = diffusion
        self._weights = np.ones([diffusion.num_timesteps])

    def weights(self):
        return self._weights


class LossAwareSampler(ScheduleSampler):
    """kept as a type so `isinstance(sampler, LossAwareSampler)` in TrainLoop stays meaningful (never instantiated)."""

    def update_with_local_losses(self, local_ts, local_losses):
        raise NotImplementedError

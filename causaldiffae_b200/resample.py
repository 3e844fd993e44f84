"""Timestep samplers (ref improved_diffusion/resample.py).  Sampling stays on the host numpy global RNG so that, given
the same np.random.seed, the drawn timesteps are bit-identical to the reference's.  LossSecondMomentResampler crashes in
the reference on numpy >= 1.24 (`np.int`); it is restated here with that one fix and pinned against the
reference run with `np.int = int` (tests/golden/make_samplers_golden.py).  Its cross-rank loss exchange is ONE padded
all_gather of (timestep, loss) pairs and ONE device->host copy per step instead of the reference's 2·B `.item()` syncs."""
from abc import ABC, abstractmethod

import numpy as np
import torch as th


def create_named_schedule_sampler(name, diffusion):
    """ref resample.py:9-21"""
    if name == "uniform":
        return UniformSampler(diffusion)
    if name == "loss-second-moment":
        return LossSecondMomentResampler(diffusion)
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
    def update_with_local_losses(self, local_ts, local_losses):
        """ref resample.py:72-103: every rank contributes its (timesteps, losses); all ranks end with the same history."""
        import torch.distributed as dist
        pair = th.stack([local_ts.detach().to(th.float64), local_losses.detach().to(th.float64)], dim=1)   # [b, 2]
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            sizes = [th.zeros(1, dtype=th.int64, device=pair.device) for _ in range(world)]
            dist.all_gather(sizes, th.tensor([pair.shape[0]], dtype=th.int64, device=pair.device))
            sizes = th.cat(sizes).cpu().tolist()
            padded = th.zeros(max(sizes), 2, dtype=th.float64, device=pair.device)
            padded[:pair.shape[0]] = pair
            parts = [th.zeros_like(padded) for _ in range(world)]
            dist.all_gather(parts, padded)
            pair = th.cat([p_[:n] for p_, n in zip(parts, sizes)], dim=0)
        host = pair.cpu().numpy()                                   # the one device->host copy of the step
        self.update_with_all_losses(host[:, 0].astype(np.int64).tolist(), host[:, 1].astype(np.float32).tolist())

    @abstractmethod
    def update_with_all_losses(self, ts, losses):
        """ts: list of int timesteps, losses: list of float losses (identical on every rank)"""


class LossSecondMomentResampler(LossAwareSampler):
    """ref resample.py:122-156: weights = sqrt(mean(loss_history^2)) mixed with a uniform floor once every timestep
    has `history_per_term` losses."""

    def __init__(self, diffusion, history_per_term=10, uniform_prob=0.001):
        self.diffusion = diffusion
        self.history_per_term = history_per_term
        self.uniform_prob = uniform_prob
        self._loss_history = np.zeros([diffusion.num_timesteps, history_per_term], dtype=np.float64)
        self._loss_counts = np.zeros([diffusion.num_timesteps], dtype=np.int64)

    def weights(self):
        if not self._warmed_up():
            return np.ones([self.diffusion.num_timesteps], dtype=np.float64)
        weights = np.sqrt(np.mean(self._loss_history ** 2, axis=-1))
        weights /= np.sum(weights)
        weights *= 1 - self.uniform_prob
        weights += self.uniform_prob / len(weights)
        return weights

    def update_with_all_losses(self, ts, losses):
        for t, loss in zip(ts, losses):             # in order: a timestep drawn twice in a batch shifts twice
            if self._loss_counts[t] == self.history_per_term:
                self._loss_history[t, :-1] = self._loss_history[t, 1:]
                self._loss_history[t, -1] = loss
            else:
                self._loss_history[t, self._loss_counts[t]] = loss
                self._loss_counts[t] += 1

    def _warmed_up(self):
        return bool((self._loss_counts == self.history_per_term).all())

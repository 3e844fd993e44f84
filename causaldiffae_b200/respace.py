"""Timestep respacing (ref improved_diffusion/respace.py).  Integer work here is bit-exact with the reference:
`space_timesteps`, `timestep_map`; the respaced float64 betas are bit-identical as well.  The per-call rebuild of
`map_tensor` from a Python list (ref respace.py:119-124) is replaced by a device-resident int64 table."""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion, ddim_coef_table  # noqa: F401  (re-exported)


def space_timesteps(num_timesteps, section_counts):
    """ref respace.py:7-61. "ddimN" -> first integer stride yielding exactly N steps; else per-section fractional
    strides rounded with Python's round()."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, kept = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """ref respace.py:65-109: keep `use_timesteps` of a base process; betas recomputed from the retained alpha-bars."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last, new_betas = 1.0, []
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def ddim_sample(self, model, *args, **kwargs):
        return super().ddim_sample(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps, self)

    def _scale_timesteps(self, t):
        return t   # done by the wrapped model, as in the reference


class _WrappedModel:
    """ref respace.py:112-124: ts -> timestep_map[ts] (and *1000/T as float when rescale_timesteps)."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, owner=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._owner = owner

    def _map_tensor(self, device, dtype):
        make = lambda: np.asarray(self.timestep_map, dtype=np.int64)  # noqa: E731
        if self._owner is not None:
            t = self._owner._dev_table("timestep_map", device, make)
        else:
            t = th.from_numpy(make()).to(device)
        return t if t.dtype == dtype else t.to(dtype)

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x, ts, **kwargs):
        new_ts = self._map_tensor(ts.device, ts.dtype)[ts.long()]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)

"""Timestep respacing with the reference's API (/root/reference/diffusion/respace.py, RS:line).

`space_timesteps` picks which of the original steps a shortened process keeps;
`SpacedDiffusion` rebuilds the betas for that subsequence and maps spaced indices back to
original timesteps before they reach the model.  The mapping table lives on the device (the
reference rebuilds and uploads it on every model call, RS:125).
"""
from __future__ import annotations

import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Set of original timesteps to keep (RS:12-62).

    section_counts: list of ints, a comma-separated string of ints (steps per equal section),
    or "ddimN" for the fixed integer stride of the DDIM paper."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept = []
    offset = 0
    for idx, count in enumerate(section_counts):
        size = base + (1 if idx < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        kept += [offset + round(pos) for pos in _strided(count, stride)]
        offset += size
    return set(kept)


def _strided(count, stride):
    """0, stride, 2*stride, ... accumulated by repeated addition exactly as RS:52-58 does (the
    float accumulation order decides how .5 cases round)."""
    pos = 0.0
    for _ in range(count):
        yield pos
        pos += stride


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process that visits only `use_timesteps` of a base process (RS:65-114)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        prev = 1.0
        betas = []
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                betas.append(1 - ac / prev)
                prev = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(betas)
        super().__init__(**kwargs)
        self._map_dev = {}

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.original_num_steps, self._map_dev)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def p_sample(self, model, *args, **kwargs):
        return super().p_sample(self._wrap_model(model), *args, **kwargs)

    def ddim_sample(self, model, *args, **kwargs):
        return super().ddim_sample(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def ddim_reverse_sample(self, model, *args, **kwargs):
        return super().ddim_reverse_sample(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def condition_score(self, cond_fn, *args, **kwargs):
        return super().condition_score(self._wrap_model(cond_fn), *args, **kwargs)

    def _scale_timesteps(self, t):
        return t  # the wrapped model does the mapping (RS:112-114)


class _WrappedModel:
    """Callable that translates spaced step indices into original timesteps (RS:117-129)."""

    def __init__(self, model, timestep_map, original_num_steps, cache=None):
        self.model = model
        self.timestep_map = timestep_map
        self.original_num_steps = original_num_steps
        self._cache = cache if cache is not None else {}

    def map_timesteps(self, ts):
        key = (ts.device.type, ts.device.index, ts.dtype)
        table = self._cache.get(key)
        if table is None:
            table = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
            self._cache[key] = table
        return table[ts]

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x, ts, **kwargs):
        return self.model(x, self.map_timesteps(ts), **kwargs)

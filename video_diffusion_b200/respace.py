"""Timestep respacing with the reference's API (`improved_diffusion/respace.py`:
`space_timesteps` :7-60, `SpacedDiffusion` :63-101, `_WrappedModel` :103-119)."""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Retained original timesteps for 'ddimN' (fixed integer stride) or per-section counts."""
    if isinstance(section_counts, str):
        if section_counts.startswith('ddim'):
            want = int(section_counts[len('ddim'):])
            for stride in range(1, num_timesteps):
                steps = range(0, num_timesteps, stride)
                if len(steps) == want:
                    return set(steps)
            raise ValueError(f'cannot create exactly {num_timesteps} steps with an integer stride')
        section_counts = [int(x) for x in section_counts.split(',')]
    base, extra = divmod(num_timesteps, len(section_counts))
    chosen, start = [], 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f'cannot divide section of {size} steps into {count}')
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        chosen += [start + round(pos) for pos in _arith(count, stride)]
        start += size
    return set(chosen)


def _arith(count, stride):
    # accumulate like the reference (cur_idx += frac_stride) so rounding ties fall the same way
    pos, out = 0.0, []
    for _ in range(count):
        out.append(pos)
        pos += stride
    return out


class SpacedDiffusion(GaussianDiffusion):
    """Diffusion over a subset of the base process' timesteps; betas are re-derived so the
    retained alphas_cumprod values are preserved."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs['betas'])
        base = GaussianDiffusion(**kwargs)
        self.timestep_map = [i for i in range(self.original_num_steps) if i in self.use_timesteps]
        kept = base.alphas_cumprod[self.timestep_map]
        prev = np.concatenate([[1.0], kept[:-1]])
        kwargs['betas'] = 1 - kept / prev
        super().__init__(**kwargs)
        self._map_dev = {}

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps, self._map_dev)

    def _eps(self, model, x, t, model_kwargs, return_attn_weights=False):
        return super()._eps(self._wrap_model(model), x, t, model_kwargs, return_attn_weights)

    def _scale_timesteps(self, t):
        return t            # the wrapped model rescales


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, cache=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._cache = cache if cache is not None else {}

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x, timesteps, **kwargs):
        key = str(timesteps.device)
        if key not in self._cache:      # the reference rebuilds this tensor on every call (:113-115)
            self._cache[key] = th.tensor(self.timestep_map, device=timesteps.device, dtype=th.long)
        # An index outside the respaced range raises IndexError in the reference (:116).  Here the lookup is clamped
        # -- a device-side assert would poison the CUDA context -- and the sampler kernels, which receive the same
        # unclamped t, record it: the next ops.check_timesteps() raises the IndexError.
        tmap = self._cache[key]
        if (self.rescale_timesteps and timesteps.is_cuda and timesteps.dtype == th.long and timesteps.dim() == 1
                and timesteps.is_contiguous()):
            # gather + rescale as one launch (the same fp32 product as below)
            from . import ops
            with th.cuda.device(timesteps.device):
                new_ts = ops.map_timesteps(timesteps, tmap, 1000.0 / self.original_num_steps)
            return self.model(x, timesteps=new_ts, **kwargs)
        new_ts = tmap[timesteps.long().clamp(0, tmap.numel() - 1)]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, timesteps=new_ts, **kwargs)

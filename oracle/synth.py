"""Deterministic synthetic weights / inputs shared by the golden generator, the tests,
smoke() and bench.py.  TEST INFRASTRUCTURE ONLY (numpy legacy RandomState: the
stream is stable across numpy/torch versions and machines).

The reference zero-initialises every ResBlock out-conv, attention proj_out, RPE-net
output layer and the final conv (unet.py:155-160, 278-279, 419, 747-748), so a
random-init model outputs exactly 0.  Parity therefore uses fully
random ("de-zeroed") weights generated here from the key -> shape spec alone.
"""
import numpy as np
import torch


def make_state_dict(spec, seed=1):
    """spec: {key: shape}.  Returns {key: fp32 tensor}; values depend only on (spec, seed)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for key in sorted(spec):
        shape = tuple(spec[key])
        z = rng.standard_normal(shape).astype(np.float32)
        if key == 'spatial_encoding':
            val = z
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            val = z * fan_in ** -0.5
        elif 'norm' in key or key.endswith('in_layers.0.weight') or key.endswith('out_layers.0.weight') \
                or key == 'out.0.weight':
            val = 1.0 + 0.1 * z if key.endswith('weight') else 0.1 * z
        elif key.endswith('in_layers.0.bias') or key.endswith('out_layers.0.bias') or key == 'out.0.bias':
            val = 0.1 * z
        else:
            val = 0.1 * z                                   # conv / linear biases
        sd[key] = torch.from_numpy(np.ascontiguousarray(val))
    return sd


def make_video(shape, seed=2, quantise=True):
    """x0 = 2*U[0,1)-1, optionally snapped to the uint8 grid."""
    rng = np.random.RandomState(seed)
    x = 2.0 * rng.random_sample(shape) - 1.0
    if quantise:
        x = np.round((x + 1) * 127.5) / 127.5 - 1
    return torch.from_numpy(x.astype(np.float32))


def make_noise(shape, seed=3):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def fingerprint(t, n=16):
    """Compact signature of a tensor: [mean, std, absmax, n strided samples]."""
    f = t.detach().float().reshape(-1)
    # fp32 linspace (kept for the committed fixtures) cannot represent indices above 2^24 exactly
    big = f.numel() > 2 ** 24
    idx = torch.linspace(0, f.numel() - 1, n, dtype=torch.float64 if big else torch.float32).long().clamp_(max=f.numel() - 1)
    return torch.cat([torch.stack([f.mean(), f.std(), f.abs().max()]), f[idx]]).numpy()

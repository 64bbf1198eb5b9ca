"""TEST INFRASTRUCTURE (oracle side) — deterministic, procedurally generated weights.

The reference ships no checkpoints and its zero_module() layers make a freshly constructed UNet
output exactly 0 (openai_model/model.py:206-208,531; openai_model/attention.py:330), so parity needs
non-trivial weights that can be regenerated anywhere without the reference and without torch's RNG
(whose stream is not guaranteed across versions).  numpy's PCG64 stream is stable, so a state dict
is a pure function of (seed, [(key, shape), ...]).

Scale rule (keeps activations at the same order of magnitude as PyTorch's default init):
  ndim >= 2 (conv / linear weight): U(-1, 1) / sqrt(fan_in)
  ndim == 1, key endswith "weight" (norm gain): 1 + 0.1 * U(-1, 1)
  ndim == 1 otherwise (bias): 0.1 * U(-1, 1)
"""
import zlib

import numpy as np
import torch


def _fan_in(shape):
    n = 1
    for s in shape[1:]:
        n *= int(s)
    return max(n, 1)


def make_tensor(key, shape, seed):
    h = zlib.crc32(key.encode()) & 0xFFFFFFFF
    rng = np.random.Generator(np.random.PCG64([seed, h]))
    u = rng.random(size=tuple(shape), dtype=np.float64) * 2.0 - 1.0
    if len(shape) >= 2:
        v = u / np.sqrt(_fan_in(shape))
    elif key.endswith("weight"):
        v = 1.0 + 0.1 * u
    else:
        v = 0.1 * u
    return torch.from_numpy(v.astype(np.float32))


def make_state_dict(key_shapes, seed):
    """key_shapes: iterable of (key, shape). Returns an ordered dict of fp32 CPU tensors."""
    return {k: make_tensor(k, tuple(s), seed) for k, s in key_shapes}


def key_shapes_of(module_or_sd):
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    return [(k, tuple(v.shape)) for k, v in sd.items() if v.dtype.is_floating_point]


def seeded_randn(shape, seed):
    """Inputs: standard normals from PCG64 (stable across platforms), fp32 CPU tensor."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal(size=tuple(shape), dtype=np.float64).astype(np.float32))

"""PRNG keys and direct samplers, bit-compatible with `jax.random`.

Mirrors `matfree/backend/prng.py:6-29` (prng_key, split, normal, uniform is
omitted, rademacher).  Keys are `uint32[2]` NumPy arrays exactly like
`jax.random.PRNGKey`; samples are generated on the GPU by `mf_probe_gen`.
`split` runs Threefry-2x32 on the host in Python integers (a handful of words).
"""

from __future__ import annotations

import numpy as np

_M = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M


def _threefry2x32(k0, k1, x0, x1):
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & _M
    x1 = (x1 + ks[1]) & _M
    rots = ((13, 15, 26, 6), (17, 29, 16, 24))
    for i in range(5):
        for r in rots[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


def prng_key(seed: int):
    """`jax.random.PRNGKey(seed)`."""
    seed = int(seed)
    return np.array([(seed >> 32) & _M, seed & _M], dtype=np.uint32)


def split(key, num: int = 2):
    """`jax.random.split` (partitionable Threefry, the JAX default since 0.5.0)."""
    k0, k1 = int(key[0]), int(key[1])
    return np.array([_threefry2x32(k0, k1, 0, j) for j in range(num)], dtype=np.uint32)


def _generate(key, shape, dtype, sampler):
    import torch

    from matfree_b200 import _lib, _device, config

    lib = _lib.load()
    tdt = _device.torch_dtype(dtype)
    shape = tuple(int(s) for s in shape)
    total = int(np.prod(shape)) if shape else 1
    out = torch.empty(max(total, 1), dtype=tdt, device=_device.device())
    # one "probe" of length total: counter = flat index
    _lib.check(lib.mf_probe_gen(out.data_ptr(), _device.mf_dtype(tdt), _lib.MF_LAYOUT_PROBE_MAJOR,
                                total, total, 0, 1, int(key[0]), int(key[1]), sampler,
                                config.prng_flags(tdt), None, _device.stream()))
    return out[:total].reshape(shape)


def normal(key, *, shape, dtype=None):
    return _generate(key, shape, dtype, 1)


def rademacher(key, *, shape, dtype=None):
    return _generate(key, shape, dtype, 0)

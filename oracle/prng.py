"""Oracle restatement of the `jax.random` stream matfree's samplers draw from.

TEST INFRASTRUCTURE (see `oracle/__init__.py`).  The arithmetic lives in the
third-party, un-pinned dependency `jax` (`/root/reference/pyproject.toml:27-30`);
matfree reaches it through `matfree/backend/prng.py:6-29`:

    prng_key(seed)            -> jax.random.PRNGKey(seed)         (prng.py:6-7)
    split(key, num)           -> jax.random.split                 (prng.py:10-11)
    normal(key, shape, dtype) -> jax.random.normal                (prng.py:14-17)
    uniform(...)              -> jax.random.uniform               (prng.py:20-23)
    rademacher(...)           -> jax.random.rademacher            (prng.py:26-29)

The published algorithm restated here is Threefry-2x32 with 20 rounds
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) used in
counter mode the way `jax/_src/prng.py` does it, in both of JAX's modes:

  * ``partitionable`` (`jax_threefry_partitionable=True`, default since JAX
    0.5.0): element with row-major flat index ``i`` of the requested shape gets
    ``TF(key; hi=i>>32, lo=i&0xffffffff)``; a 32-bit draw is ``x0 ^ x1``, a
    64-bit draw ``(x0<<32)|x1``.
  * ``legacy``: the ``N`` counters ``0..N-1`` are padded to even length and
    split into a first half (fed as ``x0``) and a second half (fed as ``x1``);
    the draws are ``concat(out0, out1)[:N]``.

Known answers are in `tests/golden/prng_kat.json`.
"""

from __future__ import annotations

import numpy as np

_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds.  `key` = (k0, k1); x0, x1 uint32 arrays."""
    k0 = np.uint32(key[0])
    k1 = np.uint32(key[1])
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    x0 = np.array(x0, dtype=np.uint32, copy=True)
    x1 = np.array(x1, dtype=np.uint32, copy=True)
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            for r in _ROT_A if i % 2 == 0 else _ROT_B:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + np.uint32(i + 1)
    return x0, x1


def prng_key(seed: int):
    """`jax.random.PRNGKey(seed)`: uint32[2] = [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2, *, mode: str = "partitionable"):
    """`jax.random.split`.  Returns uint32[num, 2]."""
    if mode == "partitionable":
        # fold-like split: new key j = TF(key; 0, j)
        lo = np.arange(num, dtype=np.uint32)
        hi = np.zeros(num, dtype=np.uint32)
        a, b = threefry2x32(key, hi, lo)
        return np.stack([a, b], axis=1)
    if mode == "legacy":
        bits = _bits32_legacy(key, 2 * num)
        return bits.reshape(num, 2)
    raise ValueError(mode)


def _bits32_legacy(key, count: int):
    ctr = np.arange(count, dtype=np.uint32)
    if count % 2:
        ctr = np.concatenate([ctr, np.zeros(1, np.uint32)])
    half = ctr.size // 2
    a, b = threefry2x32(key, ctr[:half], ctr[half:])
    return np.concatenate([a, b])[:count]


def random_bits(key, shape, *, bit_width: int = 32, mode: str = "partitionable",
                offset: int = 0):
    """Raw draws for `shape`.

    `offset` shifts the flat element index (partitionable mode only); it lets a
    caller generate rows ``p0..p1`` of the `(P, n)` sample array without the
    rows before them: ``offset = p0 * n``.
    """
    shape = tuple(int(s) for s in shape)
    size = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if mode == "partitionable":
        idx = np.arange(size, dtype=np.uint64) + np.uint64(offset)
        hi = (idx >> np.uint64(32)).astype(np.uint32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        a, b = threefry2x32(key, hi, lo)
        if bit_width == 32:
            return (a ^ b).reshape(shape)
        if bit_width == 64:
            return ((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)).reshape(shape)
        raise ValueError(bit_width)
    if mode == "legacy":
        if bit_width != 32 or offset:
            raise ValueError("legacy mode: 32-bit draws without offset only")
        return _bits32_legacy(key, size).reshape(shape)
    raise ValueError(mode)


def uniform(key, shape=(), dtype=np.float32, *, mode: str = "partitionable", offset: int = 0):
    """`jax.random.uniform(key, shape, dtype)` on [0, 1)."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        bits = random_bits(key, shape, bit_width=32, mode=mode, offset=offset)
        f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)
        return f - np.float32(1.0)
    if dtype == np.float64:
        bits = random_bits(key, shape, bit_width=64, mode=mode, offset=offset)
        f = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64)
        return f - 1.0
    raise TypeError(dtype)


def rademacher(key, shape, dtype=np.float32, *, mode: str = "partitionable",
               x64=None, offset: int = 0):
    """`jax.random.rademacher`: ``2*bernoulli(key, 0.5) - 1`` cast to `dtype`.

    ``bernoulli`` is ``uniform(key, shape, float_default) < 0.5`` so ``+1`` iff
    the top mantissa bit of the uniform is 0, i.e. iff the MSB of the draw is 0.
    With `jax_enable_x64` the comparison is done on a float64 uniform, which
    consumes a 64-bit draw whose MSB is the MSB of ``x0``.

    ``x64=None`` (default) means "the mode in which the reference can produce `dtype`": a
    float64 sample only exists under `jax_enable_x64`, so float64 -> x64 stream, float32 -> x32
    stream.  Pinned for the x32 stream by the reference's README doctest
    (`tests/golden/readme_doctest.json`).
    """
    if x64 is None:
        x64 = np.dtype(dtype) == np.float64
    if x64:
        bits = random_bits(key, shape, bit_width=64, mode=mode, offset=offset)
        neg = (bits >> np.uint64(63)).astype(np.int8)
    else:
        bits = random_bits(key, shape, bit_width=32, mode=mode, offset=offset)
        neg = (bits >> np.uint32(31)).astype(np.int8)
    return (1 - 2 * neg).astype(dtype)


# Giles' single-precision erfinv polynomial as expanded by XLA (`ErfInv32`).
_ERFINV_CENTRAL = np.array(
    [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
     -0.00125372503, -0.00417768164, 0.246640727, 1.50140941], dtype=np.float32)
_ERFINV_TAIL = np.array(
    [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
     -0.0076224613, 0.00943887047, 1.00167406, 2.83297682], dtype=np.float32)


def erf_inv_f32(x):
    """XLA's fp32 `erf_inv` expansion (Giles 2010), evaluated in fp32."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        lt = w < np.float32(5.0)
        w = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(lt, _ERFINV_CENTRAL[0], _ERFINV_TAIL[0]).astype(np.float32)
        for i in range(1, 9):
            c = np.where(lt, _ERFINV_CENTRAL[i], _ERFINV_TAIL[i]).astype(np.float32)
            p = (c + p * w).astype(np.float32)
        out = (p * x).astype(np.float32)
        out = np.where(np.abs(x) == 1, x * np.float32(np.inf), out)
    return out.astype(np.float32)


def normal(key, shape=(), dtype=np.float32, *, mode: str = "partitionable", offset: int = 0):
    """`jax.random.normal`: ``sqrt(2) * erf_inv(uniform(lo=nextafter(-1,0), hi=1))``."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
        u01 = uniform(key, shape, np.float32, mode=mode, offset=offset)
        scale = np.float32(np.float32(1.0) - lo)  # rounds to exactly 2.0f
        u = np.maximum(lo, (u01 * scale + lo).astype(np.float32))
        return (np.float32(np.sqrt(2.0)) * erf_inv_f32(u)).astype(np.float32)
    if dtype == np.float64:
        from scipy.special import erfinv

        lo = np.nextafter(-1.0, 0.0)
        u01 = uniform(key, shape, np.float64, mode=mode, offset=offset)
        u = np.maximum(lo, u01 * (1.0 - lo) + lo)
        return np.sqrt(2.0) * erfinv(u)
    raise TypeError(dtype)

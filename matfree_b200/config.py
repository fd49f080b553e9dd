"""Package-level switches that mirror the JAX configuration the reference runs under.

``jax_enable_x64`` is the only way matfree computes in float64 (`jax.config.update(
"jax_enable_x64", True)`, e.g. `tests/test_decomp/test_hessenberg_adjoint.py:71`).  Besides the
dtype it changes the probe stream: `jax.random.rademacher` is ``2*bernoulli(0.5) - 1`` and
`bernoulli` compares a uniform of the *default float dtype* with ``0.5``, so with x64 enabled
every Rademacher sample consumes a 64-bit Threefry draw whose sign bit is the MSB of ``x0``
instead of the MSB of the 32-bit draw ``x0 ^ x1`` (`matfree/backend/prng.py:26-29`; SURVEY.md
App. A.5).  This library cannot read JAX's flag, so:

* ``x64 = None`` (default, "auto"): a float64 sampler draws the x64 stream (a float64
  ``*_like`` only exists in the reference when x64 is enabled), a float32 sampler the x32 stream;
* ``update("jax_enable_x64", True / False)`` forces one stream for every dtype (True also
  covers an explicit float32 sampler inside an x64 program).
"""

from __future__ import annotations

_STATE = {"x64": None}


def update(name: str, value) -> None:
    """``matfree_b200.config.update("jax_enable_x64", True | False | None)``."""
    if name not in ("jax_enable_x64", "enable_x64", "x64"):
        raise KeyError(f"unknown configuration option {name!r}")
    if value is not None:
        value = bool(value)
    _STATE["x64"] = value


def x64_enabled(dtype=None) -> bool:
    """Whether samplers of `dtype` (a torch dtype) draw the ``jax_enable_x64`` stream."""
    if _STATE["x64"] is not None:
        return _STATE["x64"]
    import torch

    return dtype == torch.float64


def prng_flags(dtype=None) -> int:
    from matfree_b200 import _lib

    return _lib.MF_PRNG_X64_BITS if x64_enabled(dtype) else 0

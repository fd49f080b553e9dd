"""CPU oracle for the matfree SLQ hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a NumPy/SciPy (plus one small C file) restatement of the
arithmetic of pnkraemer/matfree's stochastic-Lanczos-quadrature path
(`matfree/stochtrace.py`, `matfree/funm.py`, `matfree/decomp.py`) and of the
`jax.random` Threefry-2x32 sample stream that path draws its probes from.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it, and only as the checker or as the
reported CPU baseline.  Nothing under `matfree_b200/` imports it; the product
path raises if the CUDA library is missing rather than falling back to this.

Pinning status ("Oracle"):
  * PRNG: pinned by the three Random123 Threefry-2x32 known-answer vectors and
    by documented `jax.random` outputs for key 0 in both PRNG modes
    (`tests/golden/prng_kat.json`).
  * Lanczos / quadrature numerics: pinned by the reference tests' own dense
    ground-truth identities (`tests/test_decomp/test_tridiag_sym.py`,
    `tests/test_funm/test_monte_carlo_funm_sym_logdet.py`, ...) restated in
    `tests/test_oracle_*.py`.
  * The reference itself needs JAX, which is absent from this image and from
    the wheelhouse, so no output of the reference run here exists:
    bit-level parity with the JAX implementation is UNPINNED ("parity
    unpinned" for everything beyond the KATs and identities above).
"""

from oracle import prng, ref  # noqa: F401

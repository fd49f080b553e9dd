"""FP64 tensor-core (DMMA) dense / Gram contraction against fp64 NumPy and the CUDA-core kernel.

With `jax_enable_x64` the reference's dense matvec is XLA's fp64 `dot_general`
(matfree/stochtrace.py:47-49 with tutorials/1_log_determinants.py:19-21); the bar is 1e-10
relative (BASELINE.json north_star, x64) -- the DMMA accumulates in fp64, so we ask for 1e-13.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402


def mfb():
    import matfree_b200

    return matfree_b200


@pytest.fixture(autouse=True)
def _reset_gemm_config():
    from matfree_b200 import _lib

    yield
    _lib.load().mf_gemm_config(0, 1)


def rel_err(got, want):
    return float(np.abs(got - want).max() / np.abs(want).max())


# n = 1001 (odd lda) is outside the DMMA envelope and must still be right (CUDA cores)
@pytest.mark.parametrize("n,P", [(128, 32), (200, 64), (1000, 128), (522, 256), (1000, 300), (1001, 64), (36, 32)])
def test_dense_dmma_matches_numpy(n, P):
    from matfree_b200 import _lib

    m = mfb()
    lib = _lib.load()
    A = oprng.normal(oprng.prng_key(7), (n, n), np.float64)
    A = (A + A.T) / 2
    V = oprng.normal(oprng.prng_key(2), (P, n), np.float64)
    want = (A @ V.T).T
    op = m.ops.dense(A)
    _lib.check(lib.mf_gemm_config(0, 1))
    got = op.matmat(V).cpu().numpy()
    _lib.check(lib.mf_gemm_config(0, 0))
    simt = op.matmat(V).cpu().numpy()
    assert rel_err(got, want) < 1e-13, rel_err(got, want)
    assert rel_err(simt, want) < 1e-13
    # fp64 fma chains in a different order: equal to rounding, not necessarily bitwise
    assert rel_err(got, simt) < 1e-13


@pytest.mark.parametrize("mrows,n,P", [(96, 64, 32), (300, 130, 64), (1200, 1000, 128), (513, 250, 256)])
def test_gram_dmma_matches_numpy(mrows, n, P):
    from matfree_b200 import _lib

    m = mfb()
    lib = _lib.load()
    B = oprng.normal(oprng.prng_key(5), (mrows, n), np.float64) / np.sqrt(mrows)
    V = oprng.normal(oprng.prng_key(2), (P, n), np.float64)
    want = ((B.T @ B) @ V.T).T
    op = m.ops.gram(B)
    _lib.check(lib.mf_gemm_config(0, 1))
    got = op.matmat(V).cpu().numpy()
    assert rel_err(got, want) < 1e-13, rel_err(got, want)


def test_slq_logdet_x64_dense_dmma_matches_oracle():
    """x64 SLQ log-determinant through the DMMA contraction: 1e-10 vs the fp64 oracle."""
    m = mfb()
    n, P, k = 200, 64, 12
    eig = np.linspace(1.0, 9.0, n)
    A = ref.hermitian_matrix_from_eigenvalues(eig, oprng.prng_key(3), dtype=np.float64)
    key = m.prng.prng_key(1)
    sampler = m.stochtrace.sampler_signs(np.ones(n, np.float64), num=P)
    for reortho in ("none", "full"):
        integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho=reortho))
        est = m.stochtrace.estimator_monte_carlo(integrand, sampler)
        quad = est.per_probe(m.ops.dense(A), key).cpu().numpy()
        V = oprng.rademacher(oprng.prng_key(1), (P, n), np.float64)
        oq, _ = ref.slq_batched(lambda X: X @ A.T, V, k, reortho=reortho)
        assert np.max(np.abs(quad - oq)) <= 1e-10 * np.abs(oq).max(), reortho

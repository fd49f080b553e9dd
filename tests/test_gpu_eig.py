"""eig.eigh_partial / eig.svd_partial against dense LAPACK, as in
tests/test_eig/test_eigh_partial.py:8-22 and tests/test_eig/test_svd_partial.py."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402


def mfb():
    import matfree_b200

    return matfree_b200


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_eigh_partial_equal_to_linalg_eigh(dtype):
    m = mfb()
    nrows = 10
    eigvals = np.arange(1.0, 1.0 + nrows).astype(dtype)
    A = ref.hermitian_matrix_from_eigenvalues(eigvals, oprng.prng_key(1), dtype=dtype)
    v0 = np.ones((nrows,), dtype)
    vals, vecs = m.eig.eigh_partial(m.decomp.tridiag_sym(nrows, reortho="full"))(m.ops.dense(A), v0)
    vals, vecs = vals.cpu().numpy(), vecs.cpu().numpy()
    S, U = np.linalg.eigh(A.astype(np.float64))
    tol = 1e-4 if dtype == np.float32 else 1e-9
    assert np.allclose(vals, S, rtol=tol, atol=tol)
    assert np.allclose(vecs.T @ vecs, U @ U.T, atol=tol * 10, rtol=tol * 10)
    # rows are unit Ritz vectors: A v = lambda v
    assert np.abs(A.astype(np.float64) @ vecs.T - vecs.T * vals).max() < tol * 100


@pytest.mark.parametrize("k", [8, 4, 0])
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_eigh_partial_shapes_and_oracle(k, reortho):
    m = mfb()
    nrows = 8
    A = ref.hermitian_matrix_from_eigenvalues(np.arange(1.0, 1.0 + nrows), oprng.prng_key(1), dtype=np.float64)
    v0 = oprng.normal(oprng.prng_key(2), (nrows,), np.float64)
    S, U = m.eig.eigh_partial(m.decomp.tridiag_sym(k, reortho=reortho))(m.ops.dense(A), v0)
    assert S.shape == (k,) and U.shape == (k, nrows)
    if k:
        oS, oU = ref.eigh_partial(ref.tridiag_sym(k, reortho=reortho))(lambda v: A @ v, v0)
        assert np.allclose(S.cpu().numpy(), oS, rtol=1e-8, atol=1e-8)
        sign = np.sign(np.sum(U.cpu().numpy() * oU, axis=1, keepdims=True))
        assert np.allclose(U.cpu().numpy() * sign, oU, atol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_svd_partial_equal_to_linalg_svd(dtype):
    m = mfb()
    nrows, ncols = 12, 9
    d = np.arange(ncols) + 1.0
    A = ref.asymmetric_matrix_from_singular_values(d, nrows=nrows, ncols=ncols).astype(dtype)
    v0 = oprng.normal(oprng.prng_key(1), (ncols,), dtype)
    ut, s, vt = m.eig.svd_partial(m.decomp.bidiag(ncols))(m.ops.rect(A), v0)
    ut, s, vt = ut.cpu().numpy().astype(np.float64), s.cpu().numpy(), vt.cpu().numpy().astype(np.float64)
    assert ut.shape == (ncols, nrows) and vt.shape == (ncols, ncols)
    tol = 2e-4 if dtype == np.float32 else 1e-9
    assert np.allclose(np.sort(s), np.sort(d), rtol=tol)
    assert np.abs(ut.T @ np.diag(s) @ vt - A).max() < tol * 10 * d.max()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("reortho,reortho_vjp", [("full", "match"), ("none", "match"), ("none", "none")])
@pytest.mark.parametrize("n,k", [(12, 7), (300, 24), (64, 64)])
def test_hessenberg_decomposition_nonsymmetric(dtype, reortho, reortho_vjp, n, k):
    """decomp.hessenberg on a NON-symmetric dense operator: A Q^T = Q^T H + r e_k^T, orthonormal Q,
    upper-Hessenberg H, and the same H as the oracle (tests/test_decomp/test_hessenberg.py)."""
    m = mfb()
    A = (oprng.normal(oprng.prng_key(1), (n, n), dtype) / np.sqrt(n)).astype(dtype)
    v = oprng.normal(oprng.prng_key(2), (n,), dtype)
    # the forward pass re-orthogonalises unless reortho_vjp == "none" (decomp.py:393-396,466)
    Q, H, r, c = m.decomp.hessenberg(k, reortho=reortho, reortho_vjp=reortho_vjp)(m.ops.dense(A), v)
    Q, H, r = (x.cpu().numpy().astype(np.float64) for x in (Q, H, r))
    assert Q.shape == (k, n) and H.shape == (k, k)
    tol = 2e-5 if dtype == np.float32 else 1e-11
    if reortho_vjp != "none" or k < 30:
        assert np.abs(Q @ Q.T - np.eye(k)).max() < 20 * tol
    assert np.abs(np.tril(H, -2)).max() == 0.0
    ek = np.eye(k)[:, -1]
    Ad = A.astype(np.float64)
    assert np.abs(Ad @ Q.T - Q.T @ H - np.outer(r, ek)).max() < 20 * tol
    assert np.allclose(float(c), 1 / np.linalg.norm(v.astype(np.float64)), rtol=1e-6)
    if k <= 24:
        oQ, oH, orr, oc = ref.hessenberg(k, reortho=reortho, reortho_vjp=reortho_vjp)(
            lambda x: Ad @ x, v.astype(np.float64))
        assert np.abs(H - oH).max() < 200 * tol


def test_eigh_partial_accepts_hessenberg():
    """tests/test_eig/test_eigh_partial.py:8-22 passes `decomp.hessenberg` to `eigh_partial`."""
    m = mfb()
    nrows = 10
    A = ref.hermitian_matrix_from_eigenvalues(np.arange(1.0, 1.0 + nrows), oprng.prng_key(1), dtype=np.float64)
    vals, vecs = m.eig.eigh_partial(m.decomp.hessenberg(nrows, reortho="full"))(m.ops.dense(A), np.ones(nrows))
    assert np.allclose(vals.cpu().numpy(), np.arange(1.0, 1.0 + nrows), rtol=1e-8)
    with pytest.raises(TypeError, match="Unexpected input"):
        m.decomp.hessenberg(3, reortho="partial")

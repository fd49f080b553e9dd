"""decomp.bidiag and funm.monte_carlo_funm_product_*  against the
oracle and the reference tests' identities (tests/test_decomp/test_bidiag.py,
tests/test_funm/test_monte_carlo_funm_product_logdet.py, ..._schatten_norm.py)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402


def mfb():
    import matfree_b200

    return matfree_b200


def _matrix(nrows, ncols, dtype, kind):
    if kind == "hilbert":
        a = np.arange(0, max(ncols, nrows))
        return (1.0 / (1.0 + a[:, None] + a[None, :]))[:nrows, :ncols].astype(dtype)
    n = min(nrows, ncols)
    d = np.arange(n) + 10.0
    d[4:] = 0.001
    return ref.asymmetric_matrix_from_singular_values(d, nrows=nrows, ncols=ncols).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nrows,ncols,k,kind", [(50, 49, 6, "spectrum"), (15, 13, 12, "hilbert"),
                                                (13, 15, 12, "hilbert"), (15, 15, 12, "hilbert"),
                                                (300, 128, 20, "spectrum")])
def test_bidiag_decomposition_is_satisfied(dtype, nrows, ncols, k, kind):
    m = mfb()
    A = _matrix(nrows, ncols, dtype, kind)
    v0 = oprng.normal(oprng.prng_key(1), (ncols,), dtype)
    op = m.ops.rect(A)
    assert np.allclose(op(v0).cpu().numpy(), A @ v0, rtol=1e-5, atol=1e-5)
    (U, V), B, res, ln = m.decomp.bidiag(k, materialize=True)(op, v0)
    U, V, B, res = (x.cpu().numpy().astype(np.float64) for x in (U, V, B, res))
    assert U.shape == (k, nrows) and V.shape == (k, ncols) and B.shape == (k, k)
    tol = 2e-5 if dtype == np.float32 else 1e-11
    Ad = A.astype(np.float64)
    assert np.allclose(U @ U.T, np.eye(k), atol=10 * tol)
    assert np.allclose(V @ V.T, np.eye(k), atol=10 * tol)
    em = np.eye(k)[:, -1]
    scale = np.abs(Ad).max()
    assert np.abs(Ad @ V.T - U.T @ B).max() <= 20 * tol * scale
    assert np.abs(Ad.T @ U.T - V.T @ B.T - np.outer(res, em)).max() <= 20 * tol * scale
    assert np.allclose(float(ln), 1.0 / np.linalg.norm(v0), rtol=1e-6)
    # oracle: same operation order => same B up to rounding where the spectrum is resolved
    if kind == "spectrum":
        (_, _), Bo, reso, _ = ref.bidiag(k, materialize=True)(Ad, v0.astype(np.float64))
        nsig = 4
        assert np.allclose(np.diag(B)[:nsig], np.diag(Bo)[:nsig], rtol=50 * tol)
    (d, e) = m.decomp.bidiag(k, materialize=False)(op, v0)[1]
    assert np.allclose(d.cpu().numpy(), np.diag(B)) and np.allclose(e.cpu().numpy(), np.diag(B, 1))


def test_bidiag_argument_checks():
    m = mfb()
    op = m.ops.rect(np.ones((5, 4), np.float32))
    with pytest.raises(ValueError, match="exceeds"):
        m.decomp.bidiag(5)(op, np.ones(4, np.float32))
    with pytest.raises(ValueError, match="exceeds"):
        m.decomp.bidiag(-1)(op, np.ones(4, np.float32))
    with pytest.raises(TypeError):
        m.decomp.bidiag(2)(3.0, np.ones(4, np.float32))  # neither a registered operator nor a callable
    with pytest.raises(TypeError, match="ops.rect"):
        m.decomp.bidiag(2)(m.ops.gram(np.ones((5, 4), np.float32)), np.ones(4, np.float32))
    (U, V), B, res, ln = m.decomp.bidiag(0)(op, np.ones(4, np.float32))
    assert U.shape == (0, 5) and V.shape == (0, 4) and float(res.abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_logdet_product_exact_for_full_num_matvecs(dtype):
    m = mfb()
    n = 50
    A = ref.asymmetric_matrix_from_singular_values(np.sqrt(np.arange(1.0, 1.0 + n)), nrows=n, ncols=n).astype(dtype)
    op = m.ops.rect(A)
    x = (oprng.normal(oprng.prng_key(1), (n,), dtype) + 1).astype(dtype)
    w, Q = np.linalg.eigh(A.astype(np.float64).T @ A.astype(np.float64))
    xd = x.astype(np.float64)
    tol = 1e-3 if dtype == np.float32 else 1e-9
    got = float(m.funm.monte_carlo_funm_product_logdet(m.decomp.bidiag(n - 1))(op, x))
    assert np.allclose(got, xd @ (Q @ np.diag(np.log(w)) @ Q.T) @ xd, atol=tol, rtol=tol)
    got = float(m.funm.monte_carlo_funm_product_schatten_norm(3, m.decomp.bidiag(n - 1))(op, x))
    assert np.allclose(got, xd @ (Q @ np.diag(w ** 1.5) @ Q.T) @ xd, rtol=max(tol, 1e-4))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("P", [5, 64, 150])
def test_product_logdet_estimator_matches_oracle_per_probe(dtype, P):
    """Block route (mf_probe_gen -> bidiag_blocked -> mf_bidiag_quad) against the per-probe oracle."""
    m = mfb()
    nrows, ncols, k = 80, 48, 14
    d = np.linspace(1.0, 4.0, ncols)
    A = ref.asymmetric_matrix_from_singular_values(d, nrows=nrows, ncols=ncols).astype(dtype)
    op = m.ops.rect(A)
    key = m.prng.prng_key(3)
    sampler = m.stochtrace.sampler_signs(np.ones(ncols, dtype), num=P)
    integrand = m.funm.monte_carlo_funm_product_logdet(m.decomp.bidiag(k))
    est = m.stochtrace.estimator_monte_carlo_mean_and_sem(integrand, sampler)
    vals = m.stochtrace.estimator_monte_carlo(integrand, sampler).per_probe(op, key).cpu().numpy()
    V = oprng.rademacher(oprng.prng_key(3), (P, ncols), dtype)
    ointegrand = ref.monte_carlo_funm_product_logdet(ref.bidiag(k))
    ovals = np.array([ointegrand(A.astype(np.float64), v.astype(np.float64)) for v in V])
    tol = 1e-5 if dtype == np.float32 else 1e-10
    assert np.max(np.abs(vals - ovals)) <= tol * np.abs(ovals).max()
    mean, sem = est(op, key)
    assert np.allclose(float(mean), ovals.mean(), rtol=tol)
    assert np.allclose(float(sem), ovals.std() / np.sqrt(P), rtol=1e-3)
    # single-vector integrand call == the block route's value for that probe
    one = float(integrand(op, V[0]))
    assert np.allclose(one, vals[0], rtol=10 * tol)


def test_logdet_product_estimate_is_accurate():
    """tests/test_funm/test_monte_carlo_funm_product_logdet.py:17-41 (atol = rtol = 1e-2)."""
    m = mfb()
    nrows, ncols, k = 50, 30, 20
    d = np.arange(ncols) + 1.0
    A = ref.asymmetric_matrix_from_singular_values(d, nrows=nrows, ncols=ncols).astype(np.float32)
    sampler = m.stochtrace.sampler_signs({"fx": np.ones((ncols,), np.float32)}, num=400)
    est = m.stochtrace.estimator_monte_carlo(m.funm.monte_carlo_funm_product_logdet(m.decomp.bidiag(k)), sampler)
    got = float(est(m.ops.rect(A), m.prng.prng_key(3)))
    expected = np.linalg.slogdet(A.astype(np.float64).T @ A.astype(np.float64))[1]
    assert np.allclose(got, expected, atol=1e-2, rtol=1e-2)

"""User callables ``matvec(vec, *params)`` with pytree vectors through the decompositions -- the
reference's own tests restated with torch functions in place of JAX functions (paths under
/root/reference).  The callable supplies the product; all other vector arithmetic is the CUDA
library's (C-ABI building blocks), checked against the NumPy oracle / dense ground truth."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402


def mfb():
    import matfree_b200

    return matfree_b200


def dev(x):
    return torch.as_tensor(np.asarray(x), device="cuda")


def matvec_pytree(s, p):
    # tests/test_decomp/test_tridiag_sym.py:15-17
    [(x,)] = s
    return [(p @ x,)]


# tests/test_decomp/test_tridiag_sym.py:7-38
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_tridiag_full_rank_reconstruction_pytree_and_params(reortho):
    m = mfb()
    ndim = 12
    eigvals = np.arange(1.0, 2.0, 1 / ndim)
    matrix = ref.hermitian_matrix_from_eigenvalues(eigvals, oprng.prng_key(1))
    vector = np.flip(np.arange(1.0, 1.0 + ndim)).copy()
    algorithm = m.decomp.tridiag_sym(ndim, reortho=reortho, materialize=True)
    Q_pytree, T, *_ = algorithm(matvec_pytree, [(vector,)], dev(matrix))
    [(Q,)] = Q_pytree
    Q, T = Q.cpu().numpy(), T.cpu().numpy()
    tol = 1e-5 if reortho == "full" else 1e-1
    assert np.allclose(Q.T @ T @ Q, matrix, atol=tol, rtol=tol)
    ref.assert_columns_orthonormal(Q)
    ref.assert_columns_orthonormal(Q.T)


# tests/test_decomp/test_tridiag_sym.py:43-65, plus agreement with the registered-operator route
@pytest.mark.parametrize("num_matvecs", [1, 5, 11])
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_tridiag_mid_rank_decomposition_callable_equals_registered(num_matvecs, reortho):
    m = mfb()
    ndim = 12
    eigvals = np.arange(1.0, 2.0, 1 / ndim)
    matrix = ref.hermitian_matrix_from_eigenvalues(eigvals, oprng.prng_key(1))
    vector = np.flip(np.arange(1.0, 1.0 + ndim)).copy()
    algorithm = m.decomp.tridiag_sym(num_matvecs, reortho=reortho, materialize=True)
    Q_pytree, T, q_pytree, _n = algorithm(matvec_pytree, [(vector,)], dev(matrix))
    [(Q,)] = Q_pytree
    [(q,)] = q_pytree
    Q, T, q = Q.cpu().numpy(), T.cpu().numpy(), q.cpu().numpy()
    e_K = np.eye(num_matvecs)[-1]
    ref.assert_allclose(matrix @ Q.T - Q.T @ T - np.outer(q, e_K), np.zeros((ndim, num_matvecs)))
    Q2, T2, q2, n2 = algorithm(m.ops.dense(matrix), vector)
    assert np.allclose(T, T2.cpu().numpy(), atol=1e-11)
    assert np.allclose(Q, Q2.cpu().numpy(), atol=1e-10)
    assert np.allclose(float(_n), float(n2), rtol=1e-14)
    Qo, To, qo, co = ref.tridiag_sym(num_matvecs, reortho=reortho)(lambda x: matrix @ x, vector)
    assert np.allclose(T, To, atol=1e-11) and np.allclose(float(_n), co, rtol=1e-14)


def test_registered_operator_rejects_params():
    m = mfb()
    A = np.eye(4, dtype=np.float32)
    with pytest.raises(TypeError, match="only supported for callables"):
        m.decomp.tridiag_sym(2)(m.ops.dense(A), np.ones(4, np.float32), dev(A))
    with pytest.raises(TypeError, match="registered operator or a callable"):
        m.decomp.tridiag_sym(2)(3.0, np.ones(4, np.float32))


# tests/test_decomp/test_hessenberg.py:7-40
@pytest.mark.parametrize("num_matvecs", [0, 5, 9])
@pytest.mark.parametrize("reortho", ["none", "full"])
def test_hessenberg_decomposition_is_satisfied_pytree(num_matvecs, reortho):
    m = mfb()
    nrows = 10
    A = oprng.normal(oprng.prng_key(1), (nrows, nrows), np.float64)
    v = oprng.normal(oprng.prng_key(2), (nrows,), np.float64)
    algorithm = m.decomp.hessenberg(num_matvecs, reortho=reortho)
    Q_pytree, H, r_pytree, c = algorithm(matvec_pytree, [(v,)], dev(A))
    [(Q,)] = Q_pytree
    [(r,)] = r_pytree
    assert tuple(Q.shape) == (num_matvecs, nrows) and tuple(H.shape) == (num_matvecs, num_matvecs)
    assert tuple(r.shape) == (nrows,) and tuple(c.shape) == ()
    Q, H, r, c = Q.cpu().numpy(), H.cpu().numpy(), r.cpu().numpy(), float(c)
    e = np.eye(num_matvecs)
    if num_matvecs:
        ref.assert_allclose(A @ Q.T - Q.T @ H - np.outer(r, e[-1]), np.zeros((nrows, num_matvecs)))
        ref.assert_allclose(Q @ Q.T - e, np.zeros_like(e))
        ref.assert_allclose(Q.T @ e[0], c * v)
        oQ, oH, orr, oc = ref.hessenberg(num_matvecs, reortho=reortho)(lambda x: A @ x, v)
        assert np.allclose(H, oH, atol=1e-11) and np.allclose(r, orr, atol=1e-10)


# tests/test_funm/test_funm_lanczos_sym.py:7-37
@pytest.mark.parametrize("dense_funm", ["eigh", "schur"])
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_funm_lanczos_sym_matches_eigh_implementation(dense_funm, reortho):
    m = mfb()
    n = 11
    v = oprng.normal(oprng.prng_key(2), (n,), np.float64)
    eigvals = np.linspace(0.01, 0.99, n)
    matrix = ref.hermitian_matrix_from_eigenvalues(eigvals, oprng.prng_key(1))
    lam, vecs = np.linalg.eigh(matrix)
    expected = (vecs @ np.diag(np.sin(lam)) @ vecs.T) @ v
    fun = torch.sin if dense_funm == "eigh" else np.sin
    df = (m.funm.dense_funm_sym_eigh if dense_funm == "eigh" else m.funm.dense_funm_schur)(fun)
    lanczos = m.decomp.tridiag_sym(6, materialize=True, reortho=reortho)
    matfun_vec = m.funm.funm_lanczos_sym(df, lanczos)
    [(received,)] = matfun_vec(matvec_pytree, [(v,)], dev(matrix))
    assert np.allclose(expected, received.cpu().numpy(), atol=1e-6)


# tests/test_funm/test_funm_arnoldi.py:11-37
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_funm_arnoldi_matches_dense_expm(reortho):
    import scipy.linalg

    m = mfb()
    n = 11
    matrix = oprng.normal(oprng.prng_key(1), (n, n), np.float64)
    v = oprng.normal(oprng.prng_key(2), (n,), np.float64)
    expected = scipy.linalg.expm(matrix) @ v
    arnoldi = m.decomp.hessenberg((n * 3) // 4, reortho=reortho)
    matfun_vec = m.funm.funm_arnoldi(m.funm.dense_funm_pade_exp(), arnoldi)
    [(received,)] = matfun_vec(matvec_pytree, [(v,)], dev(matrix))
    assert np.allclose(expected, received.cpu().numpy(), rtol=1e-1, atol=1e-1)
    # full depth is exact, and a registered (non-symmetric dense) operator takes the kernel route
    full = m.funm.funm_arnoldi(m.funm.dense_funm_pade_exp(), m.decomp.hessenberg(n, reortho="full"))
    got = full(m.ops.dense(matrix), v).cpu().numpy()
    assert np.allclose(got, expected, rtol=1e-8, atol=1e-8)
    got_s = m.funm.funm_arnoldi(m.funm.dense_funm_schur(np.exp), m.decomp.hessenberg(n, reortho="full"))(
        m.ops.dense(matrix), v).cpu().numpy()
    assert np.allclose(got_s, expected, rtol=1e-6, atol=1e-6)


# tests/test_eig/test_eig_partial.py:7-24
def test_eig_partial_equal_to_linalg_eig():
    m = mfb()
    nrows = 7
    A = np.triu(np.arange(1.0, 1.0 + nrows**2).reshape(nrows, nrows))
    v0 = np.ones(nrows)
    alg = m.eig.eig_partial(m.decomp.hessenberg(nrows, reortho="full"))
    vals, vecs = alg(lambda v, p: p @ v, v0, dev(A))
    S, U = np.linalg.eig(A)
    vals, vecs = vals.cpu().numpy(), vecs.cpu().numpy()
    assert np.allclose(np.sort(vals.real), np.sort(S.real)) and np.abs(vals.imag).max() < 1e-9
    assert np.allclose(vecs.T @ vecs, U @ U.T, atol=1e-5, rtol=1e-5)


# tests/test_eig/test_eig_partial.py:27-62
@pytest.mark.parametrize("num_matvecs", [0, 2, 3])
def test_eig_partial_shapes_lists_tuples(num_matvecs):
    m = mfb()
    nrows = 10
    K = dev(np.arange(1.0, 10.0).reshape(3, 3))
    v0 = np.ones((nrows, nrows))  # tensor-valued input

    def Av(v, stencil):
        [(x,)] = v
        y = torch.nn.functional.conv2d(x[None, None], stencil.flip(0, 1)[None, None], padding=1)
        return [(y[0, 0],)]

    vals, [(vecs,)] = m.eig.eig_partial(m.decomp.hessenberg(num_matvecs, reortho="none"))(Av, [(v0,)], K)
    assert tuple(vecs.shape) == (num_matvecs, nrows, nrows) and tuple(vals.shape) == (num_matvecs,)


# tests/test_funm/test_monte_carlo_funm_sym_logdet.py:16-38 (dict-valued vectors, callable matvec);
# the reference runs it with x64 disabled, i.e. in float32
def test_logdet_spd_dict_vectors_through_the_estimator():
    m = mfb()
    n, nsig, k = 200, 30, 10
    keyA, key = oprng.split(oprng.prng_key(1))
    d = (np.arange(n) / n + 1.0).astype(np.float32)
    d[nsig:] = 0.001
    A = ref.hermitian_matrix_from_eigenvalues(d, keyA)
    At = dev(A)

    def matvec(x):
        return {"fx": At @ x["fx"]}

    sampler = m.stochtrace.sampler_normal({"fx": np.ones((n,), dtype=np.float32)}, num=10)
    samples = sampler(key)
    assert isinstance(samples, dict) and tuple(samples["fx"].shape) == (10, n)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, materialize=True))
    received = float(m.stochtrace.estimator_monte_carlo(integrand, sampler)(matvec, key))
    expected = np.linalg.slogdet(A.astype(np.float64))[1]
    assert np.allclose(received, expected, atol=1e-2, rtol=1e-2)
    # the oracle on the same key (same probes): lambda_min = 1e-3, so log amplifies fp32 roundoff
    owant = ref.estimator_monte_carlo(ref.monte_carlo_funm_sym_logdet(ref.tridiag_sym(k)),
                                      ref.sampler_normal(n, num=10, dtype=np.float32))(lambda v: A @ v, key)
    assert np.allclose(received, owant, rtol=1e-4)
    # the registered operator on the same key: same probes, same estimate (fused kernel chain)
    flat_sampler = m.stochtrace.sampler_normal(np.ones((n,), dtype=np.float32), num=10)
    fused = float(m.stochtrace.estimator_monte_carlo(integrand, flat_sampler)(m.ops.dense(A), key))
    assert np.allclose(received, fused, rtol=1e-4)


# tests/test_decomp/test_bidiag.py (callable matvec: the transpose comes from its VJP)
def test_bidiag_with_a_callable_and_params():
    m = mfb()
    nrows, ncols, k = 9, 6, 4
    A = ref.asymmetric_matrix_from_singular_values(np.linspace(1.0, 3.0, ncols), nrows=nrows, ncols=ncols)
    v0 = oprng.normal(oprng.prng_key(1), (ncols,), np.float64)
    (U, V), B, res, c = m.decomp.bidiag(k)(lambda v, p: p @ v, v0, dev(A))
    (U2, V2), B2, res2, c2 = m.decomp.bidiag(k)(m.ops.rect(A), v0)
    assert np.allclose(B.cpu().numpy(), B2.cpu().numpy(), atol=1e-11)
    assert np.allclose(U.cpu().numpy(), U2.cpu().numpy(), atol=1e-10)
    (Uo, Vo), Bo, reso, co = ref.bidiag(k)(A, v0)
    assert np.allclose(B.cpu().numpy(), Bo, atol=1e-10)


def test_hutchinson_integrands_with_pytree_vectors_and_params():
    """stochtrace.py:836-914 on a dict-valued vector and a parametrised callable (generic route)."""
    m = mfb()
    n = 6
    A = np.arange(1.0, 1.0 + n * n).reshape(n, n) / n
    sampler = m.stochtrace.sampler_signs({"a": np.ones((2,), np.float64), "b": np.ones((2, 2), np.float64)}, num=300)

    def matvec(x, p):
        flat = torch.cat([x["a"], x["b"].reshape(-1)])
        y = p @ flat
        return {"a": y[:2], "b": y[2:].reshape(2, 2)}

    key = m.prng.prng_key(4)
    est = m.stochtrace.estimator_monte_carlo(m.stochtrace.monte_carlo_trace_and_diagonal(), sampler)
    got = est(matvec, key, dev(A))
    V = oprng.rademacher(oprng.prng_key(4), (300, n), np.float64)
    want_tr = np.einsum("pi,pi->p", V, V @ A.T).mean()
    want_diag = (V * (V @ A.T)).mean(axis=0)
    assert np.allclose(float(got["trace"]), want_tr, rtol=1e-12)
    assert tuple(got["diagonal"]["b"].shape) == (2, 2)
    assert np.allclose(got["diagonal"]["a"].cpu().numpy(), want_diag[:2], rtol=1e-12)
    assert np.allclose(got["diagonal"]["b"].cpu().numpy().ravel(), want_diag[2:], rtol=1e-12)

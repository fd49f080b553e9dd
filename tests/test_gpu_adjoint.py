"""The Lanczos / Arnoldi adjoints on the GPU (`matfree_b200.adjoint`; reference:
/root/reference/matfree/decomp.py:295-348,480-600) against the NumPy oracle (oracle/adjoint.py,
itself pinned by finite differences in tests/test_oracle_adjoint.py) and against central finite
differences of the CUDA forward pass, in fp64 ("x64": 1e-8), restating
/root/reference/tests/test_decomp/test_tridiag_sym_adjoint.py:7-49 and
test_hessenberg_adjoint.py:5-113 with `torch.autograd` in place of `jax.vjp`."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import adjoint as oadj  # noqa: E402
from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402


def mfb():
    import matfree_b200

    return matfree_b200


def dev(x, grad=False):
    t = torch.as_tensor(np.asarray(x), device="cuda")
    return t.requires_grad_(grad)


def _sym_matrix(n, seed):
    key_eig, key_mat = oprng.split(oprng.prng_key(seed))
    eigvals = oprng.uniform(key_eig, (n,), np.float64) + 1.0
    return ref.hermitian_matrix_from_eigenvalues(eigvals, key_mat)


# tests/test_decomp/test_tridiag_sym_adjoint.py:7-49 (n = 10, k = 4, seeds 1..3)
@pytest.mark.parametrize("kind", ["dense", "csr", "callable"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tridiag_none_adjoint_matches_oracle(kind, seed):
    import scipy.sparse as sp

    m = mfb()
    n, k = 10, 4
    A = _sym_matrix(n, seed)
    v = oprng.normal(oprng.prng_key(1), (n,), np.float64)
    rng = np.random.default_rng(seed)
    dxs, dal, dbe = rng.standard_normal((k + 1, n)), rng.standard_normal(k), rng.standard_normal(k)
    xs, al, be, nrm = oadj.tridiag_forward_cache(A, v, k)
    (gv, gA), _ = oadj.tridiag_adjoint(A, initvec_norm=nrm, alphas=al, betas=be, xs=xs,
                                       dalphas=dal, dbetas=dbe, dxs=dxs)
    vt = dev(v, True)
    alg = m.decomp.tridiag_sym(k, reortho="none", materialize=False)
    if kind == "dense":
        At = dev(A, True)
        Q, (d, e), res, c = alg(m.ops.dense(At), vt)
    elif kind == "csr":
        S = sp.csr_matrix(A)
        data = dev(S.data, True)
        Q, (d, e), res, c = alg(m.ops.csr(S.indptr.astype(np.int32), S.indices.astype(np.int32), data), vt)
    else:
        At = dev(A, True)
        Q, (d, e), res, c = alg(lambda x, p: p @ x, vt, At)
    # forward values == the oracle's
    assert np.allclose(Q.detach().cpu().numpy(), xs[:-1], atol=1e-10)
    assert np.allclose(d.detach().cpu().numpy(), al, atol=1e-10)
    # the outputs are (Q, diag, offdiag, residual = b x_last, 1/|v|): build the same scalar
    # sum(dxs * xs) + dal . al + dbe . be the oracle differentiates
    x_last = res / torch.linalg.vector_norm(res)
    b_last = torch.linalg.vector_norm(res)
    loss = (dev(dxs[:-1]) * Q).sum() + (dev(dxs[-1]) * x_last).sum() + dev(dal) @ d \
        + dev(dbe[:-1]) @ e + float(dbe[-1]) * b_last
    loss.backward()
    assert np.allclose(vt.grad.cpu().numpy(), gv, rtol=1e-8, atol=1e-9)
    if kind == "csr":
        want = np.asarray([gA[r, c_] for r in range(n) for c_ in S.indices[S.indptr[r]:S.indptr[r + 1]]])
        assert np.allclose(data.grad.cpu().numpy(), want, rtol=1e-8, atol=1e-9)
    else:
        assert np.allclose(At.grad.cpu().numpy(), gA, rtol=1e-8, atol=1e-9)


# tests/test_decomp/test_hessenberg_adjoint.py:5-38 and :75-113
@pytest.mark.parametrize("kind", ["dense", "callable"])
@pytest.mark.parametrize("reortho", ["none", "full"])
@pytest.mark.parametrize("n,k", [(3, 2), (10, 4), (15, 10)])
def test_hessenberg_adjoint_matches_oracle(kind, reortho, n, k):
    m = mfb()
    A = oprng.normal(oprng.prng_key(1), (n, n), np.float64)
    v = oprng.normal(oprng.prng_key(2), (n,), np.float64)
    rng = np.random.default_rng(k)
    dQ, dH, dr, dc = rng.standard_normal((n, k)), np.triu(rng.standard_normal((k, k)), -1), rng.standard_normal(n), rng.standard_normal()
    oQ, oH, orr, oc = ref._hessenberg_forward(lambda x: A @ x, k, v, reortho="full")
    gv, gA = oadj.hessenberg_adjoint(A, Q=oQ, H=oH, r=orr, c=oc, dQ=dQ, dH=dH, dr=dr, dc=dc, reortho=reortho)
    vt, At = dev(v, True), dev(A, True)
    alg = m.decomp.hessenberg(k, reortho=reortho)
    Q, H, r, c = alg(m.ops.dense(At), vt) if kind == "dense" else alg(lambda x, p: p @ x, vt, At)
    assert np.allclose(H.detach().cpu().numpy(), oH, atol=1e-10)
    loss = (dev(dQ.T) * Q).sum() + (dev(dH) * H).sum() + dev(dr) @ r + float(dc) * c
    loss.backward()
    assert np.allclose(vt.grad.cpu().numpy(), gv, rtol=1e-7, atol=1e-8)
    assert np.allclose(At.grad.cpu().numpy(), gA, rtol=1e-7, atol=1e-8)


def test_hessenberg_adjoint_k_zero_raises():
    # tests/test_decomp/test_hessenberg_adjoint.py:43-62
    m = mfb()
    A = oprng.normal(oprng.prng_key(1), (3, 3), np.float64)
    vt, At = dev(np.ones(3), True), dev(A, True)
    Q, H, r, c = m.decomp.hessenberg(0, reortho="full")(m.ops.dense(At), vt)
    with pytest.raises(ValueError, match="= 0"):
        (r.sum() + c).backward()


@pytest.mark.parametrize("reortho", ["none", "full"])
def test_slq_logdet_gradient_on_a_large_csr_operator_matches_finite_differences(reortho):
    """d/d(shift) and d/d(data) of the SLQ quadratic form v^T log(A) v on a 10^4-row CSR operator
    (2-D Laplacian + shift I): autograd through the adjoint kernels against central finite
    differences of the CUDA forward pass (x64, 1e-8 relative)."""
    from matfree_b200 import workloads

    m = mfb()
    shape, k = (100, 100), 12
    n = shape[0] * shape[1]
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, dtype="float64", device="cuda")
    v = dev(oprng.normal(oprng.prng_key(3), (n,), np.float64))
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho=reortho))
    rng = np.random.default_rng(0)
    direction = dev(rng.standard_normal(d.shape[0]) * 0.1)
    # keep the perturbed operator symmetric: perturb entry (r, c) and (c, r) alike
    rows = torch.repeat_interleave(torch.arange(n, device="cuda"), (ip[1:] - ip[:-1]).long())
    key_fwd = rows * n + ix.long()
    key_bwd = ix.long() * n + rows
    order_f, order_b = torch.argsort(key_fwd), torch.argsort(key_bwd)
    sym_dir = direction.clone()
    sym_dir[order_b] = sym_dir[order_b] + direction[order_f]

    def value(data):
        return integrand(m.ops.csr(ip, ix, data), v)

    data = d.clone().requires_grad_(True)
    q = value(data)
    q.backward()
    analytic = float(data.grad @ sym_dir)
    eps = 1e-6
    with torch.no_grad():
        fd = float((value(d + eps * sym_dir) - value(d - eps * sym_dir)) / (2 * eps))
    assert np.isclose(analytic, fd, rtol=1e-6, atol=1e-8 * abs(float(q))), (analytic, fd)
    # ... and the no-grad value is the fused-kernel value
    with torch.no_grad():
        assert np.isclose(float(value(d)), float(q), rtol=1e-10)


def test_estimator_gradient_through_slq_mean():
    """`jax.grad` of an SLQ log-determinant estimate with respect to a kernel hyper-parameter
    (tutorials/8_gaussian_logpdf.py:45-66), as `loss.backward()`: exact for full-depth Lanczos."""
    m = mfb()
    n = 12
    B = oprng.normal(oprng.prng_key(5), (n, n), np.float64)
    K0 = dev(B @ B.T / n)
    theta = torch.tensor(0.7, dtype=torch.float64, device="cuda", requires_grad=True)
    eye = torch.eye(n, dtype=torch.float64, device="cuda")

    def matvec(x, scale):
        return (scale * K0 + eye) @ x

    sampler = m.stochtrace.sampler_signs(np.ones(n), num=8)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(n, reortho="full"))
    est = m.stochtrace.estimator_monte_carlo(integrand, sampler)(matvec, m.prng.prng_key(1), theta)
    est.backward()
    # full depth: every quadratic form is exact, so d/dtheta mean_p v^T log(theta K0 + I) v
    # = mean_p v^T (theta K0 + I)^{-1} K0 v
    V = oprng.rademacher(oprng.prng_key(1), (8, n), np.float64)
    K = K0.cpu().numpy()
    M = np.linalg.solve(0.7 * K + np.eye(n), K)
    want = np.mean([v @ M @ v for v in V])
    assert np.isclose(float(theta.grad), want, rtol=1e-7)

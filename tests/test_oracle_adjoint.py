"""Pin the oracle's restatement of the Lanczos / Arnoldi adjoints (oracle/adjoint.py) by central
finite differences of the oracle's forward passes, in fp64.  The reference's own tests compare the
custom VJP with JAX autodiff of the forward pass for random cotangents
(/root/reference/tests/test_decomp/test_tridiag_sym_adjoint.py:7-49,
test_hessenberg_adjoint.py:5-38,75-113); without an autodiff host the directional derivative of
the forward pass plays that role."""

import numpy as np
import pytest

from oracle import adjoint, prng, ref


def _sym_matrix(n, seed):
    # tests/test_decomp/test_tridiag_sym_adjoint.py:14-17
    key_eig, key_mat = prng.split(prng.prng_key(seed))
    eigvals = prng.uniform(key_eig, (n,), np.float64) + 1.0
    return ref.hermitian_matrix_from_eigenvalues(eigvals, key_mat)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tridiag_adjoint_matches_finite_differences(seed):
    n, k = 10, 4
    A = _sym_matrix(n, seed)
    v = prng.normal(prng.prng_key(1), (n,), np.float64)
    rng = np.random.default_rng(seed)
    dxs, dal, dbe = rng.standard_normal((k + 1, n)), rng.standard_normal(k), rng.standard_normal(k)

    def forward(vv, AA):
        xs, al, be, _ = adjoint.tridiag_forward_cache(AA, vv, k)
        return np.sum(dxs * xs) + dal @ al + dbe @ be

    xs, al, be, nrm = adjoint.tridiag_forward_cache(A, v, k)
    (gv, gA), _ = adjoint.tridiag_adjoint(A, initvec_norm=nrm, alphas=al, betas=be, xs=xs,
                                          dalphas=dal, dbetas=dbe, dxs=dxs)
    eps = 1e-6
    for _ in range(3):
        dv = rng.standard_normal(n)
        dA = rng.standard_normal((n, n))
        dA = dA + dA.T  # stay symmetric: Lanczos assumes it
        fd = (forward(v + eps * dv, A + eps * dA) - forward(v - eps * dv, A - eps * dA)) / (2 * eps)
        assert np.isclose(gv @ dv + np.sum(gA * dA), fd, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("reortho", ["none", "full"])
@pytest.mark.parametrize("n,k", [(3, 2), (10, 4), (15, 10)])
def test_hessenberg_adjoint_matches_finite_differences(reortho, n, k):
    A = prng.normal(prng.prng_key(1), (n, n), np.float64)
    v = prng.normal(prng.prng_key(2), (n,), np.float64)
    rng = np.random.default_rng(k)
    dQ, dH, dr, dc = rng.standard_normal((n, k)), rng.standard_normal((k, k)), rng.standard_normal(n), rng.standard_normal()
    dH = np.triu(dH, -1)  # H is upper Hessenberg: cotangents of structural zeros are irrelevant

    def fwd(vv, AA):
        Q, H, r, c = ref._hessenberg_forward(lambda x: AA @ x, k, vv, reortho="full")
        return Q, H, r, c

    def scalar(vv, AA):
        Q, H, r, c = fwd(vv, AA)
        return np.sum(dQ * Q) + np.sum(dH * H) + dr @ r + dc * c

    Q, H, r, c = fwd(v, A)
    gv, gA = adjoint.hessenberg_adjoint(A, Q=Q, H=H, r=r, c=c, dQ=dQ, dH=dH, dr=dr, dc=dc, reortho=reortho)
    eps = 1e-6
    for _ in range(3):
        dv, dA = rng.standard_normal(n), rng.standard_normal((n, n))
        fd = (scalar(v + eps * dv, A + eps * dA) - scalar(v - eps * dv, A - eps * dA)) / (2 * eps)
        assert np.isclose(gv @ dv + np.sum(gA * dA), fd, rtol=2e-5, atol=1e-6), (reortho, n, k)


def test_hessenberg_adjoint_k_zero_raises():
    # tests/test_decomp/test_hessenberg_adjoint.py:43-62
    with pytest.raises(ValueError, match="= 0"):
        adjoint.hessenberg_adjoint(np.eye(3), Q=np.zeros((3, 0)), H=np.zeros((0, 0)), r=np.ones(3), c=1.0,
                                   dQ=np.zeros((3, 0)), dH=np.zeros((0, 0)), dr=np.ones(3), dc=0.0, reortho="full")

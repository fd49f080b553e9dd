"""The C port of the oracle (CPU baseline) agrees with the NumPy oracle."""

import numpy as np
import scipy.sparse as sp

from oracle import port, prng, ref


def _lap(shape, shift=1.0):
    def lap1(m):
        return sp.diags([-np.ones(m - 1), 2 * np.ones(m), -np.ones(m - 1)], [-1, 0, 1])

    A = sp.kron(lap1(shape[0]), sp.eye(shape[1])) + sp.kron(sp.eye(shape[0]), lap1(shape[1]))
    return (A + shift * sp.eye(shape[0] * shape[1])).tocsr().astype(np.float32)


def test_port_probes_bit_exact():
    key = prng.prng_key(1)
    X = port.probes(501, 7, 3, key)
    want = prng.rademacher(key, (7, 501), np.float32, offset=3 * 501)
    assert np.array_equal(X, want.T)


def test_port_slq_matches_numpy_oracle():
    A = _lap((20, 23))
    n = A.shape[0]
    key = prng.prng_key(1)
    P, k = 12, 10
    quad, al, be = port.csr_logdet_quadforms(A.indptr, A.indices, A.data, key, 0, P, k, return_coeffs=True)
    V = prng.rademacher(key, (P, n), np.float32)
    oq, otheta = ref.slq_batched(lambda X: (A @ X.T).T, V, k, reortho="none")
    assert np.allclose(quad, oq, rtol=2e-5)
    a, b, _ = ref.lanczos_none_batched(lambda X: (A @ X.T).T, V, k)
    assert np.allclose(al, a, rtol=1e-4, atol=1e-5) and np.allclose(be, b, rtol=1e-4, atol=1e-5)
    # probe offsets give slices of the same sample array
    q2 = port.csr_logdet_quadforms(A.indptr, A.indices, A.data, key, 5, 4, k)
    assert np.allclose(q2, quad[5:9], rtol=1e-6)


def test_port_trace_matches():
    A = _lap((9, 11))
    key = prng.prng_key(2)
    got = port.csr_trace_samples(A.indptr, A.indices, A.data, key, 0, 9)
    V = prng.rademacher(key, (9, 99), np.float32)
    want = np.einsum("pn,pn->p", V, (A @ V.T).T)
    assert np.allclose(got, want, rtol=1e-5)
    assert port.num_threads() >= 1

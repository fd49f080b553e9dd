"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances follow BASELINE.json's north_star: probes bit-exact; Ritz values,
trace and log-determinant estimates within 1e-5 relative (fp32) / 1e-10 (fp64).
Test cases restate the reference's own tests (cited per test, paths under
/root/reference) with registered operators in place of JAX callables.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402

RTOL = {np.float32: 1e-5, np.float64: 1e-10}


def mfb():
    import matfree_b200

    return matfree_b200


def to_np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def lap_scipy(shape, shift, dtype):
    import scipy.sparse as sp

    from matfree_b200 import workloads

    ip, ix, d = workloads.laplacian_csr(shape, shift=shift, dtype=np.dtype(dtype).name)
    n = int(np.prod(shape))
    return sp.csr_matrix((d.numpy(), ix.numpy(), ip.numpy()), shape=(n, n))


def spd_dense(n, dtype, seed=5, lo=1.0, hi=9.0):
    eig = np.linspace(lo, hi, n).astype(dtype)
    return ref.hermitian_matrix_from_eigenvalues(eig, oprng.prng_key(seed), dtype=dtype)


# ------------------------------------------------------------------ K1 probes


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,num", [(5, 7), (1000, 33), (4099, 256)])
def test_sampler_signs_bit_exact(dtype, n, num):
    m = mfb()
    key = m.prng.prng_key(1)
    x_like = np.ones(n, dtype=dtype)
    got = to_np(m.stochtrace.sampler_signs(x_like, num=num)(key))
    want = oprng.rademacher(oprng.prng_key(1), (num, n), dtype)
    assert got.dtype == dtype and got.shape == (num, n)
    assert np.array_equal(got, want)


def test_sampler_signs_x64_stream_and_config_switch():
    """`jax_enable_x64` changes the Rademacher stream (64-bit draw, sign = MSB of x0;
    /root/reference/matfree/backend/prng.py:26-29): float64 samplers draw it by
    default, `config.update("jax_enable_x64", ...)` forces either stream for every dtype -- in the
    materialised sampler AND in the fused estimator (`mf_estimate`, probes never materialised)."""
    m = mfb()
    key, okey = m.prng.prng_key(1), oprng.prng_key(1)
    n, num = 515, 37
    want64 = oprng.rademacher(okey, (num, n), np.float64, x64=True)
    want32 = oprng.rademacher(okey, (num, n), np.float64, x64=False)
    assert not np.array_equal(want64, want32)
    assert np.array_equal(to_np(m.stochtrace.sampler_signs(np.ones(n, np.float64), num=num)(key)), want64)
    As = lap_scipy((5, 103), 1.0, np.float64)
    op = m.ops.csr_from_scipy(As)
    trace = m.stochtrace.monte_carlo_trace()
    try:
        for flag, dt, want in ((True, np.float32, want64), (False, np.float64, want32), (None, np.float64, want64),
                               (None, np.float32, want32)):
            m.config.update("jax_enable_x64", flag)
            s = m.stochtrace.sampler_signs(np.ones(n, dt), num=num)
            assert np.array_equal(to_np(s(key)), want.astype(dt)), (flag, dt)
            assert np.array_equal(to_np(m.prng.rademacher(key, shape=(num, n), dtype=dt)), want.astype(dt))
            # fused route: per-probe v^T A v from in-kernel probes == the same from the oracle's probes
            opd = op if dt == np.float64 else m.ops.csr_from_scipy(As.astype(np.float32))
            got = to_np(m.stochtrace.estimator_monte_carlo(trace, s).per_probe(opd, key))
            ref_vals = np.einsum("pn,pn->p", want, (As @ want.T).T)
            assert np.allclose(got, ref_vals, rtol=1e-6 if dt == np.float32 else 1e-13), (flag, dt)
    finally:
        m.config.update("jax_enable_x64", None)


@pytest.mark.parametrize("n,num", [(5, 7), (1000, 33)])
def test_sampler_normal_fp32(n, num):
    m = mfb()
    key = m.prng.prng_key(3)
    got = to_np(m.stochtrace.sampler_normal(np.ones(n, np.float32), num=num)(key))
    want = oprng.normal(oprng.prng_key(3), (num, n), np.float32)
    # uniform bits are exact; erf_inv's log1p differs by <= 2 ulp between libms
    assert np.allclose(got, want, rtol=4e-6, atol=1e-7)
    assert np.mean(got == want) > 0.5


def test_sampler_normal_fp64():
    m = mfb()
    got = to_np(m.stochtrace.sampler_normal(np.ones(300, np.float64), num=9)(m.prng.prng_key(3)))
    want = oprng.normal(oprng.prng_key(3), (9, 300), np.float64)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-14)


def test_blocked_probes_match_reference_layout_and_offsets():
    """Tile generation (blocked layout, probe offset, 64-bit counters) == slices of the (P, n) array."""
    from matfree_b200 import _device, _lib

    lib = _lib.load()
    n, ld, p0, npb = 777, 32, 5_000_000, 19  # p0 * n > 2^32: high counter word is live
    out = torch.empty((n, ld), dtype=torch.float32, device="cuda")
    _lib.check(lib.mf_probe_gen(out.data_ptr(), 0, _lib.MF_LAYOUT_BLOCKED, n, ld, p0, npb, 0, 1, 0, 0,
                                None, _device.stream()))
    got = to_np(out)
    want = oprng.rademacher(oprng.prng_key(1), (npb, n), np.float32, offset=p0 * n)
    assert np.array_equal(got[:, :npb], want.T)
    assert np.all(got[:, npb:] == 0)


def test_prng_split_and_direct_samplers():
    m = mfb()
    assert np.array_equal(m.prng.split(m.prng.prng_key(0)), oprng.split(oprng.prng_key(0)))
    got = to_np(m.prng.normal(m.prng.prng_key(0), shape=(), dtype=np.float32))
    assert np.allclose(got, 1.6226422, rtol=1e-6)  # documented jax.random.normal(key(0))
    r = to_np(m.prng.rademacher(m.prng.prng_key(1), shape=(3, 5), dtype=np.float32))
    assert np.array_equal(r, oprng.rademacher(oprng.prng_key(1), (3, 5), np.float32))


# ------------------------------------------------------------------ K2 operators


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("P", [1, 3, 64, 300])
def test_matmat_csr(dtype, P):
    m = mfb()
    A = lap_scipy((13, 17), 0.5, dtype)
    op = m.ops.csr_from_scipy(A)
    V = oprng.normal(oprng.prng_key(2), (P, A.shape[0]), dtype)
    got = to_np(op.matmat(V))
    want = (A @ V.T).T
    assert np.allclose(got, want, rtol=50 * np.finfo(dtype).eps, atol=50 * np.finfo(dtype).eps)
    one = to_np(op(V[0]))
    assert np.allclose(one, want[0], rtol=50 * np.finfo(dtype).eps, atol=50 * np.finfo(dtype).eps)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,P", [(37, 5), (200, 64), (130, 256)])
def test_matmat_dense_and_gram(dtype, n, P):
    m = mfb()
    A = spd_dense(n, dtype)
    V = oprng.normal(oprng.prng_key(2), (P, n), dtype)
    tol = 200 * np.finfo(dtype).eps
    got = to_np(m.ops.dense(A).matmat(V))
    want = (A.astype(np.float64) @ V.T.astype(np.float64)).T
    assert np.allclose(got, want, rtol=tol, atol=tol * np.abs(want).max())
    B = oprng.normal(oprng.prng_key(4), (n + 11, n), dtype)
    gotg = to_np(m.ops.gram(B).matmat(V))
    wantg = ((B.T.astype(np.float64) @ B.astype(np.float64)) @ V.T.astype(np.float64)).T
    assert np.allclose(gotg, wantg, rtol=tol, atol=tol * np.abs(wantg).max())


def test_callable_takes_the_generic_route_and_the_fused_entry_points_refuse_it():
    """A callable of device tensors is accepted by the decompositions (the product is the
    callable's, the recurrence the library's); the fused kernel chain needs the operator's buffers
    and says so; anything else is a TypeError.  There is no CPU route either way."""
    m = mfb()
    A = spd_dense(8, np.float32)
    At = torch.as_tensor(A, device="cuda")
    Q, T, r, c = m.decomp.tridiag_sym(3)(lambda v: At @ v, np.ones(8, np.float32))
    Q2, T2, r2, c2 = m.decomp.tridiag_sym(3)(m.ops.dense(A), np.ones(8, np.float32))
    assert np.allclose(to_np(T), to_np(T2), atol=1e-5)
    with pytest.raises(TypeError, match="registered operator"):
        m.decomp.tridiag_sym(3)("not a matvec", np.ones(8, np.float32))
    from matfree_b200 import _generic

    with pytest.raises(TypeError, match="registered"):
        _generic.CallableOperator(lambda v: v, 8, torch.float32)._struct()


# ------------------------------------------------------------------ decomp.tridiag_sym


# tests/test_decomp/test_tridiag_sym.py:7-38
@pytest.mark.parametrize("reortho", ["full", "none"])
@pytest.mark.parametrize("kind", ["dense", "csr"])
def test_tridiag_full_rank_reconstructs(reortho, kind):
    m = mfb()
    n = 12
    A = spd_dense(n, np.float64, lo=1.0, hi=12.0)
    if kind == "csr":
        import scipy.sparse as sp

        op = m.ops.csr_from_scipy(sp.csr_matrix(A))
    else:
        op = m.ops.dense(A)
    v = oprng.normal(oprng.prng_key(2), (n,), np.float64)
    Q, T, res, c = m.decomp.tridiag_sym(n, reortho=reortho)(op, v)
    Q, T, c = to_np(Q), to_np(T), to_np(c)
    tol = 1e-5 if reortho == "full" else 1e-1
    assert np.allclose(Q @ Q.T, np.eye(n), atol=tol)
    assert np.allclose(Q.T @ T @ Q, A, atol=tol * n)
    # /root/reference/matfree/decomp.py:142 vs :177 -- "full" returns |v| (1/(1/|v|)), "none" 1/|v|
    assert np.allclose(c, np.linalg.norm(v) if reortho == "full" else 1 / np.linalg.norm(v))


# tests/test_decomp/test_tridiag_sym.py:43-65
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("reortho", ["full", "none"])
@pytest.mark.parametrize("k", [1, 5, 11])
def test_tridiag_decomposition_identity_and_oracle(dtype, reortho, k):
    m = mfb()
    n = 12
    A = spd_dense(n, dtype, lo=1.0, hi=12.0)
    v = oprng.normal(oprng.prng_key(3), (n,), dtype)
    Q, T, q, c = m.decomp.tridiag_sym(k, reortho=reortho)(m.ops.dense(A), v)
    Q, T, q = to_np(Q), to_np(T), to_np(q)
    assert Q.shape == (k, n) and T.shape == (k, k) and q.shape == (n,)
    e_K = np.eye(k, dtype=dtype)[-1]
    ref.assert_allclose(A @ Q.T - Q.T @ T - np.outer(q, e_K), np.zeros((n, k), dtype=dtype))
    ref.assert_allclose(Q @ Q.T, np.eye(k, dtype=dtype))
    Qo, To, qo, co = ref.tridiag_sym(k, reortho=reortho)(lambda x: A @ x, v)
    tol = 2e-4 if dtype == np.float32 else 1e-9
    assert np.allclose(T, To, rtol=tol, atol=tol)
    assert np.allclose(Q, Qo, atol=20 * tol)
    assert np.allclose(q, qo, atol=20 * tol * np.abs(qo).max())
    assert np.allclose(to_np(c), co, rtol=tol)


# tests/test_decomp/test_consistency.py:27-63
@pytest.mark.parametrize("reortho", ["full", "none"])
@pytest.mark.parametrize("k", [6, 13, 0])
def test_shapes(reortho, k):
    m = mfb()
    n = 13
    A = spd_dense(n, np.float32)
    Q, T, r, c = m.decomp.tridiag_sym(k, reortho=reortho)(m.ops.dense(A), np.ones(n, np.float32))
    assert tuple(Q.shape) == (k, n) and tuple(T.shape) == (k, k)
    assert tuple(r.shape) == (n,) and tuple(c.shape) == ()
    d, e = m.decomp.tridiag_sym(k, reortho=reortho, materialize=False)(m.ops.dense(A), np.ones(n, np.float32))[1]
    assert tuple(d.shape) == (k,) and tuple(e.shape) == (max(k - 1, 0),)


@pytest.mark.parametrize("reortho", ["full", "none"])
@pytest.mark.parametrize("k", [-1, 14])
def test_num_matvecs_out_of_range(reortho, k):
    m = mfb()
    A = spd_dense(13, np.float32)
    with pytest.raises(ValueError, match="exceeds"):
        m.decomp.tridiag_sym(k, reortho=reortho)(m.ops.dense(A), np.ones(13, np.float32))


def test_unknown_reortho():
    with pytest.raises(ValueError, match="unsupported"):
        mfb().decomp.tridiag_sym(3, reortho="partial")


# ------------------------------------------------------------------ funm / SLQ


# tests/test_funm/test_monte_carlo_funm_sym_logdet.py:41-67
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_logdet_exact_for_full_depth(reortho):
    m = mfb()
    n = 50
    eig = np.arange(1.0, 1.0 + n)
    A = ref.hermitian_matrix_from_eigenvalues(eig, oprng.prng_key(1), dtype=np.float64)
    x = oprng.normal(oprng.prng_key(1), (n,), np.float64) + 10
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(n - 1, reortho=reortho))
    got = to_np(integrand(m.ops.dense(A), x))
    lam, U = np.linalg.eigh(A)
    want = x @ (U @ np.diag(np.log(lam)) @ U.T) @ x
    assert np.allclose(got, want)


# tests/test_funm/test_monte_carlo_funm_sym_logdet.py:16-38 (the reference's seeds and dtype)
def test_logdet_spd_reference_case():
    m = mfb()
    n, nsig, k = 200, 30, 10
    key_A, key = m.prng.split(m.prng.prng_key(1))
    d = np.arange(n, dtype=np.float32) / np.float32(n) + np.float32(1.0)
    d[nsig:] = 0.001
    A = ref.hermitian_matrix_from_eigenvalues(d, key_A)
    sampler = m.stochtrace.sampler_normal(np.ones(n, np.float32), num=10)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, materialize=True))
    estimate = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    got = float(estimate(m.ops.dense(A), key))
    want = np.linalg.slogdet(A)[1]
    assert np.allclose(got, want, atol=1e-2, rtol=1e-2)
    # and against the oracle run on the same key
    osampler = ref.sampler_normal(n, num=10, dtype=np.float32)
    ointegrand = ref.monte_carlo_funm_sym_logdet(ref.tridiag_sym(k))
    owant = ref.estimator_monte_carlo(ointegrand, osampler)(lambda v: A @ v, key)
    assert np.allclose(got, owant, rtol=1e-4)  # lambda_min = 1e-3: log amplifies fp32 roundoff


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("reortho", ["none", "full"])
@pytest.mark.parametrize("kind", ["csr", "dense", "gram"])
def test_slq_per_probe_matches_oracle(dtype, reortho, kind):
    """Per-probe quadratic forms, Ritz values and the estimate vs the oracle on the same key."""
    m = mfb()
    P, k = 40, 12
    if kind == "csr":
        As = lap_scipy((24, 24), 1.0, dtype)
        n = As.shape[0]
        op = m.ops.csr_from_scipy(As)
        matmat = lambda X: (As @ X.T).T  # noqa: E731
    elif kind == "dense":
        n = 300
        A = spd_dense(n, dtype)
        op = m.ops.dense(A)
        matmat = lambda X: X @ A.T  # noqa: E731
    else:
        n = 160
        B = (oprng.normal(oprng.prng_key(4), (400, n), dtype) / np.sqrt(400)).astype(dtype)
        op = m.ops.gram(B)
        matmat = lambda X: (X @ B.T) @ B  # noqa: E731
    key = m.prng.prng_key(1)
    sampler = m.stochtrace.sampler_signs(np.ones(n, dtype), num=P)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho=reortho))
    estimate = m.stochtrace.estimator_monte_carlo_mean_and_sem(integrand, sampler)
    plain = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    quad, alphas, betas, lens = plain.per_probe(op, key, return_coeffs=True)
    V = oprng.rademacher(oprng.prng_key(1), (P, n), dtype)
    oq, otheta = ref.slq_batched(matmat, V, k, reortho=reortho)
    rtol = RTOL[dtype]
    # Ritz values
    nodes, _ = m.funm.ritz_blocked(alphas[0], betas[0], P)
    theta = to_np(nodes).T
    assert np.allclose(theta, otheta, rtol=rtol, atol=0), np.max(np.abs(theta - otheta) / np.abs(otheta))
    # per-probe quadratic forms: compare on the scale of the estimate
    scale = np.abs(oq).mean()
    assert np.max(np.abs(to_np(quad) - oq)) <= 3 * rtol * scale
    mean, sem = estimate(op, key)
    assert np.allclose(float(mean), oq.astype(np.float64).mean(), rtol=rtol)
    assert np.allclose(float(sem), oq.astype(np.float64).std() / np.sqrt(P), rtol=1e-3)
    assert np.allclose(float(plain(op, key)), float(mean))
    assert np.allclose(to_np(lens)[0, :P], np.sqrt(n), rtol=1e-6)


def test_slq_tiled_equals_untiled_and_partial_tiles():
    """Probe tiling (ld) must not change any per-probe value: probes are independent."""
    m = mfb()
    As = lap_scipy((20, 20), 1.0, np.float32)
    op = m.ops.csr_from_scipy(As)
    P, k = 70, 8
    sampler = m.stochtrace.sampler_signs(np.ones(400, np.float32), num=P)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho="none"))
    est = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    key = m.prng.prng_key(9)
    a = to_np(est.per_probe(op, key))
    b = to_np(est.per_probe(op, key, tile=16))
    c = to_np(est.per_probe(op, key, tile=4))
    assert a.shape == (P,)
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_slq_custom_matfun_goes_through_nodes_and_weights():
    m = mfb()
    A = spd_dense(120, np.float64)
    op = m.ops.dense(A)
    sampler = m.stochtrace.sampler_signs(np.ones(120, np.float64), num=6)
    tri = m.decomp.tridiag_sym(10, reortho="full")
    custom = m.funm.monte_carlo_funm_sym(m.funm.dense_funm_sym_eigh(lambda x: x * x + 1.0), tri)
    got = float(m.stochtrace.estimator_monte_carlo(custom, sampler)(op, m.prng.prng_key(2)))
    V = oprng.rademacher(oprng.prng_key(2), (6, 120), np.float64)
    oq, _ = ref.slq_batched(lambda X: X @ A.T, V, 10, reortho="full", matfun=lambda x: x * x + 1.0)
    assert np.allclose(got, oq.mean(), rtol=1e-10)
    # north_star spelling of the integrand factories
    assert m.funm.integrand_funm_sym_logdet is m.funm.monte_carlo_funm_sym_logdet
    assert m.funm.integrand_funm_sym is m.funm.monte_carlo_funm_sym


# tests/test_funm/test_funm_lanczos_sym.py:7-37
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_funm_lanczos_sym(reortho):
    m = mfb()
    n, k = 11, 6
    eig = np.arange(1.0, 1.0 + n) / n
    A = ref.hermitian_matrix_from_eigenvalues(eig, oprng.prng_key(1), dtype=np.float64)
    v = oprng.normal(oprng.prng_key(2), (n,), np.float64)
    lam, U = np.linalg.eigh(A)
    want = U @ (np.sin(lam) * (U.T @ v))
    fun = m.funm.funm_lanczos_sym(m.funm.dense_funm_sym_eigh(np.sin), m.decomp.tridiag_sym(k, reortho=reortho))
    got = to_np(fun(m.ops.dense(A), v))
    assert np.allclose(got, want, atol=1e-6)
    ofun = ref.funm_lanczos_sym(ref.dense_funm_sym_eigh(np.sin), ref.tridiag_sym(k, reortho=reortho))
    assert np.allclose(got, ofun(lambda x: A @ x, v), atol=1e-10)


# ------------------------------------------------------------------ Hutchinson


# tests/test_stochtrace/test_monte_carlo/test_trace.py:7-37 (real case)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_hutchinson_trace(dtype):
    m = mfb()
    As = lap_scipy((16, 16), 1.0, dtype)
    op = m.ops.csr_from_scipy(As)
    P = 500
    sampler = m.stochtrace.sampler_signs(np.ones(256, dtype), num=P)
    est = m.stochtrace.estimator_monte_carlo_mean_and_sem(m.stochtrace.monte_carlo_trace(), sampler)
    mean, sem = est(op, m.prng.prng_key(1))
    V = oprng.rademacher(oprng.prng_key(1), (P, 256), dtype)
    per = np.einsum("pn,pn->p", V, (As @ V.T).T).astype(np.float64)
    assert np.allclose(float(mean), per.mean(), rtol=RTOL[dtype])
    assert np.allclose(float(sem), per.std() / np.sqrt(P), rtol=1e-4)
    assert np.allclose(float(mean), As.diagonal().sum(), rtol=1e-2)


def test_generic_integrand_and_callable_still_work():
    """A user integrand / callable goes through the reference's generic route (sample, map, mean)."""
    m = mfb()
    J = torch.as_tensor(oprng.normal(oprng.prng_key(3), (4, 4), np.float32), device="cuda")
    sampler = m.stochtrace.sampler_normal(np.ones(4, np.float32), num=2000)
    est = m.stochtrace.estimator_monte_carlo(m.stochtrace.monte_carlo_trace(), sampler)
    got = float(est(lambda v: J @ v, m.prng.prng_key(1)))
    assert np.allclose(got, float(torch.trace(J)), atol=0.3)


# ------------------------------------------------------------------ size-independent properties


def test_full_size_properties_laplacian_2d_1024():
    """1M-row Laplacian: Rademacher init length is sqrt(n) exactly; the SLQ estimate agrees with
    the closed-form log-determinant within the Monte-Carlo error; Hutchinson trace is exact
    (zero variance would need a diagonal matrix; here within 3 sem)."""
    m = mfb()
    from matfree_b200 import workloads

    shape = (1024, 1024)
    n = shape[0] * shape[1]
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, device="cuda")
    op = m.ops.csr(ip, ix, d)
    P, k = 64, 30
    sampler = m.stochtrace.sampler_signs(np.ones(n, np.float32), num=P)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho="none"))
    est = m.stochtrace.estimator_monte_carlo_mean_and_sem(integrand, sampler)
    mean, sem = est(op, m.prng.prng_key(1))
    want = workloads.laplacian_logdet(shape, 1.0)
    assert abs(float(mean) - want) <= 4 * float(sem) + 1e-4 * abs(want)
    plain = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    _, _, _, lens = plain.per_probe(op, m.prng.prng_key(1), return_coeffs=True)
    assert np.all(to_np(lens)[0, :P] == np.float32(1024.0))
    tr = m.stochtrace.estimator_monte_carlo_mean_and_sem(m.stochtrace.monte_carlo_trace(), sampler)
    tmean, tsem = tr(op, m.prng.prng_key(1))
    assert abs(float(tmean) - 5.0 * n) <= 4 * float(tsem) + 1e-5 * n


# ------------------------------------------------------------------ irregular CSR (long rows)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("P", [1, 3, 64, 300])
def test_matmat_csr_long_rows(dtype, P):
    """Rows far above the 128-non-zero threshold (one of them spanning several 512-non-zero
    segments, one empty row) take the load-balanced route: result equals SciPy's, and it is
    deterministic run to run."""
    import scipy.sparse as sp

    m = mfb()
    rng = np.random.default_rng(0)
    n = 3000
    B = sp.random(n, n, density=0.002, random_state=rng, format="lil", dtype=np.float64)
    B[5, :] = rng.standard_normal(n)            # 3000 non-zeros: 6 segments
    B[17, ::3] = 1.5                            # 1000 non-zeros: 2 segments
    B[40, :300] = -2.0                          # 300 non-zeros: one segment
    B[41, :] = 0.0                              # empty row
    A = B.tocsr().astype(dtype)
    A.sort_indices()
    op = m.ops.csr_from_scipy(A)
    assert op.max_row_nnz == n
    V = oprng.normal(oprng.prng_key(2), (P, n), dtype)
    got = to_np(op.matmat(V))
    want = (A.astype(np.float64) @ V.T.astype(np.float64)).T
    tol = 200 * np.finfo(dtype).eps
    assert np.allclose(got, want, rtol=tol, atol=tol * np.abs(want).max())
    assert np.array_equal(got, to_np(op.matmat(V)))
    # Lanczos on top of it (alpha is formed by the separate dot kernel on this route)
    S = (A + A.T + sp.identity(n) * 80.0).tocsr().astype(dtype)
    S.sort_indices()
    sop = m.ops.csr_from_scipy(S)
    v = oprng.normal(oprng.prng_key(3), (n,), dtype)
    _, (d1, e1), _, _ = m.decomp.tridiag_sym(8, reortho="none", materialize=False)(sop, v)
    od, oe, _ = ref.lanczos_none_batched(lambda X: (S @ X.T).T, v[None, :], 8)
    rt = 2e-5 if dtype == np.float32 else 1e-10
    assert np.allclose(to_np(d1), od[0], rtol=rt, atol=rt)
    assert np.allclose(to_np(e1), oe[0][:7], rtol=rt, atol=rt)

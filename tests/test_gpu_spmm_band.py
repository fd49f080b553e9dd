"""The band route of the CSR product (csrc/spmm_strip.cu: register window over adjacent diagonals,
TMA bulk copies of the CSR metadata) against SciPy and, bit for bit, against the row-group gather
kernel it replaces on stencil matrices (`mf_spmm_config` switches between the two)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def mfb():
    import matfree_b200

    return matfree_b200


def scipy_csr(ip, ix, d, n):
    import scipy.sparse as sp

    return sp.csr_matrix((d.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))


@pytest.fixture()
def spmm_knobs():
    from matfree_b200 import _lib

    lib = _lib.load()
    yield lib
    lib.mf_spmm_config(2, 64, 2, 3)  # the library defaults


@pytest.mark.parametrize("dtype,ld", [("float32", 256), ("float32", 128), ("float64", 256), ("float64", 64)])
@pytest.mark.parametrize("shape", [(37, 70), (5, 103), (64, 256), (48, 64), (96, 128), (9, 10, 33), (3, 40, 64), (700,)])
def test_band_kernel_matches_scipy_and_gather_kernel_bitwise(spmm_knobs, dtype, ld, shape):
    from matfree_b200 import workloads

    m = mfb()
    lib = spmm_knobs
    n = int(np.prod(shape))
    ip, ix, d = workloads.laplacian_csr(shape, shift=0.5, dtype=dtype, device="cuda")
    # distinct values per entry, so that a wrong value <-> column pairing cannot cancel
    gen = torch.Generator(device="cuda").manual_seed(3)
    d = d * (1.0 + 0.25 * torch.rand(d.shape, generator=gen, device="cuda", dtype=d.dtype))
    op = m.ops.csr(ip, ix, d)
    X = torch.randn((n, ld), generator=gen, device="cuda", dtype=d.dtype)
    results = {}
    # 0: row-group gather kernel, 1: band kernel (register window), 2: TMA-staged band kernels,
    # 3: + every 7-diagonal matrix, 4: + the strip walk whatever the size, 5: without the strip walk
    for band, rows, pfd, minb in ((0, 64, 2, 4), (1, 64, 2, 4), (1, 32, 0, 3), (1, 128, -2, 4), (2, 64, 2, 3), (3, 64, 2, 3),
                                  (4, 64, 2, 3), (5, 64, 2, 3)):
        lib.mf_spmm_config(band, rows, pfd, minb)
        results[(band, rows)] = op.matmat_blocked(X)
    want = scipy_csr(ip, ix, d, n) @ X.cpu().numpy()
    tol = 1e-5 if dtype == "float32" else 1e-13
    base = results[(0, 64)]
    assert np.allclose(base.cpu().numpy(), want, rtol=tol, atol=tol * 10)
    for key, W in results.items():
        assert torch.equal(W, base), key  # the FMA order per row is the CSR order on every route


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_band_kernel_fused_alpha_dot_in_slq(spmm_knobs, dtype):
    """The Lanczos alpha (`decomp.py:288`) is reduced inside the product kernel: per-probe SLQ
    values with the band kernel == with the gather kernel == the oracle."""
    from matfree_b200 import workloads
    from oracle import prng as oprng
    from oracle import ref

    m = mfb()
    lib = spmm_knobs
    shape, P, k = (40, 64), 130, 9
    n = shape[0] * shape[1]
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, dtype=np.dtype(dtype).name, device="cuda")
    op = m.ops.csr(ip, ix, d)
    key = m.prng.prng_key(1)
    sampler = m.stochtrace.sampler_signs(np.ones(n, dtype), num=P)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho="none"))
    est = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    vals = {}
    tol = 1e-5 if dtype == np.float32 else 1e-10
    for band, tile in ((0, 128), (1, 128), (0, 256), (2, 256), (4, 256), (5, 256)):  # 4: strip walk (40 lines of 64)
        lib.mf_spmm_config(band, 64, 2, 4)
        vals[band] = est.per_probe(op, key, tile=tile).cpu().numpy()
        assert np.allclose(vals[0], vals[band], rtol=tol * 0.1, atol=0), (band, tile)
    A = scipy_csr(ip, ix, d, n)
    V = oprng.rademacher(oprng.prng_key(1), (P, n), dtype)
    oq, _ = ref.slq_batched(lambda X: (A @ X.T).T, V, k, reortho="none")
    assert np.max(np.abs(vals[1] - oq)) <= 3 * tol * np.abs(oq).mean()


def test_band_kernel_irregular_matrix_falls_back_row_by_row(spmm_knobs):
    """avg <= 5 non-zeros per row but no band structure at all: every strip takes the gather path."""
    import scipy.sparse as sp

    m = mfb()
    rng = np.random.default_rng(0)
    n, ld = 5000, 128
    A = sp.random(n, n, density=4.0 / n, random_state=rng, format="csr", dtype=np.float32)
    A.sort_indices()
    op = m.ops.csr_from_scipy(A)
    X = torch.randn((n, ld), device="cuda")
    spmm_knobs.mf_spmm_config(1, 64, 2, 4)
    W1 = op.matmat_blocked(X)
    spmm_knobs.mf_spmm_config(0, 64, 2, 4)
    W0 = op.matmat_blocked(X)
    assert torch.equal(W0, W1)
    assert np.allclose(W1.cpu().numpy(), A @ X.cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_blocked_row_order_for_wide_3d_stencils_is_bit_identical(spmm_knobs):
    """A 3-D stencil whose planes (bandwidth = 256^2 rows of 1 KB) exceed what L2 keeps between
    their uses is walked in a blocked row order (mf_operator_t::csr_bandwidth, csrc/spmm_csr.cu);
    a narrow tile of the same operator is walked in ascending order.  Row arithmetic is the same,
    so the common columns agree bit for bit -- and with SciPy to rounding."""
    from matfree_b200 import workloads

    m = mfb()
    shape = (6, 256, 256)
    n = int(np.prod(shape))
    ip, ix, d = workloads.laplacian_csr(shape, shift=0.5, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(5)
    d = d * (1.0 + 0.25 * torch.rand(d.shape, generator=gen, device="cuda"))
    op = m.ops.csr(ip, ix, d)
    assert op.bandwidth == 256 * 256
    X = torch.randn((n, 256), generator=gen, device="cuda")
    W = op.matmat_blocked(X)                      # 2 * 64 MB planes: blocked order
    Xn = X[:, :32].contiguous()
    Wn = op.matmat_blocked(Xn)                    # 2 * 8 MB planes: ascending order
    assert torch.equal(W[:, :32], Wn)
    assert op.num_diagonals == 7 and op.line_stride == 256
    spmm_knobs.mf_spmm_config(5, 64, 2, 3)        # the TMA-staged 7-diagonal kernel, blocked order
    assert torch.equal(op.matmat_blocked(X), W)
    spmm_knobs.mf_spmm_config(4, 64, 2, 3)        # the strip walk (one item = one plane's strip)
    assert torch.equal(op.matmat_blocked(X), W)
    spmm_knobs.mf_spmm_config(0, 64, 2, 3)        # the gather kernel, forced
    assert torch.equal(op.matmat_blocked(X), W)
    rows = torch.tensor([0, 1, 255, 256, 65535, 65536, 65537, n // 2 + 3, n - 65537, n - 1], device="cuda")
    A = scipy_csr(ip, ix, d, n)
    want = A[rows.cpu().numpy()] @ X.cpu().numpy()
    assert np.allclose(W[rows].cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    # the fused alpha dot rides on the same order: SLQ on the blocked order == narrow tiles
    sampler = m.stochtrace.sampler_signs(np.broadcast_to(np.float32(1), (n,)), num=256)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(4, reortho="none"))
    est = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    a = est.per_probe(op, m.prng.prng_key(2), tile=256)
    b = est.per_probe(op, m.prng.prng_key(2), tile=32)
    assert torch.allclose(a, b, rtol=1e-6)


def test_tma_kernel_partial_bands_and_diagonal_count_hint(spmm_knobs):
    """The TMA-staged kernel at ld = 256 on matrices that are only partly bands: chunks whose rows
    leave the five diagonals take its in-kernel gather path, x-boundary rows (a missing entry) stay
    on the band path with a zero coefficient.  `ops.csr` counts the diagonals
    (mf_operator_t::csr_num_diagonals): with the count known the kernel is chosen for true 5-diagonal
    matrices only; with it forced to "unknown" the average row length decides and an irregular matrix
    runs through the TMA kernel's fallback.  All routes agree bit for bit with the row-group kernel."""
    import scipy.sparse as sp
    from matfree_b200 import workloads

    m = mfb()
    gen = torch.Generator(device="cuda").manual_seed(11)
    ld = 256
    # (a) 2-D Laplacian with non-constant coefficients: 5 diagonals
    shape = (48, 64)
    n = int(np.prod(shape))
    ip, ix, d = workloads.laplacian_csr(shape, shift=0.5, device="cuda")
    d = d * (1.0 + 0.25 * torch.rand(d.shape, generator=gen, device="cuda"))
    op = m.ops.csr(ip, ix, d)
    assert op.num_diagonals == 5 and op.bandwidth == 64
    X = torch.randn((n, ld), generator=gen, device="cuda")
    spmm_knobs.mf_spmm_config(0, 64, 2, 3)
    W0 = op.matmat_blocked(X)
    for cfg in (2, 4):  # 4: the strip walk (48 lines of 64 rows)
        spmm_knobs.mf_spmm_config(cfg, 64, 2, 3)
        assert torch.equal(op.matmat_blocked(X), W0)
    assert np.allclose(W0.cpu().numpy(), scipy_csr(ip, ix, d, n) @ X.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # (b) the same matrix with some entries moved off the diagonals (rows 100..139: column + 7)
    A = scipy_csr(ip, ix, d, n).tolil()
    for r in range(100, 140):
        c = (r + 64) % n
        v = A[r, c]
        A[r, c] = 0.0
        A[r, (c + 7) % n] = v if v != 0 else 0.125
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    op2 = m.ops.csr_from_scipy(A)
    assert op2.num_diagonals == 255
    for forced, cfg in ((255, 2), (0, 2), (0, 4)):  # vetoed -> row-group kernel; unknown -> TMA kernels with per-chunk fallback
        op2.num_diagonals = forced
        spmm_knobs.mf_spmm_config(cfg, 64, 2, 3)
        W2 = op2.matmat_blocked(X)
        spmm_knobs.mf_spmm_config(0, 64, 2, 3)
        assert torch.equal(op2.matmat_blocked(X), W2)
    assert np.allclose(W2.cpu().numpy(), A @ X.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # (c) no band structure at all, 4.5 entries per row on average
    rng = np.random.default_rng(3)
    B = sp.random(n, n, density=4.5 / n, random_state=rng, format="csr", dtype=np.float32)
    B.sort_indices()
    op3 = m.ops.csr_from_scipy(B)
    assert op3.num_diagonals == 255
    for forced in (255, 0):
        op3.num_diagonals = forced
        spmm_knobs.mf_spmm_config(2, 64, 2, 3)
        W3 = op3.matmat_blocked(X)
        spmm_knobs.mf_spmm_config(0, 64, 2, 3)
        assert torch.equal(op3.matmat_blocked(X), W3)
    assert np.allclose(W3.cpu().numpy(), B @ X.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # (d) the 3-D stencil counts 7
    ip3, ix3, d3 = workloads.laplacian_csr((5, 8, 16), shift=0.5, device="cuda")
    assert m.ops.csr(ip3, ix3, d3).num_diagonals == 7


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_single_vector_row_thread_kernel_matches_wider_tiles_bitwise(dtype):
    """ld = 1 (a single vector, BASELINE config 4) takes the one-thread-per-row kernel: the same
    CSR-order FMAs as the tiled kernels, so column 0 of a wider tile has the same bits."""
    import scipy.sparse as sp
    from matfree_b200 import workloads

    m = mfb()
    gen = torch.Generator(device="cuda").manual_seed(2)
    tdt = torch.float32 if dtype == "float32" else torch.float64
    ip, ix, d = workloads.laplacian_csr((9, 10, 33), shift=0.5, dtype=dtype, device="cuda")
    d = d * (1.0 + 0.25 * torch.rand(d.shape, generator=gen, device="cuda", dtype=tdt))
    ops_ = [m.ops.csr(ip, ix, d)]
    rng = np.random.default_rng(4)
    B = sp.random(3000, 3000, density=20.0 / 3000, random_state=rng, format="csr", dtype=np.dtype(dtype))
    B.sort_indices()  # rows of 5 .. 40 entries: the loop path
    ops_.append(m.ops.csr_from_scipy(B))
    for op in ops_:
        X = torch.randn((op.n, 4), generator=gen, device="cuda", dtype=tdt)
        W4 = op.matmat_blocked(X)
        W1 = op.matmat_blocked(X[:, :1].contiguous())
        assert torch.equal(W1[:, 0], W4[:, 0])
    want = B @ X[:, :1].cpu().numpy()
    tol = 1e-5 if dtype == "float32" else 1e-12
    assert np.allclose(W1.cpu().numpy(), want, rtol=tol, atol=tol)

"""Row-sharded CSR operator and decompositions on one GPU (world size 1: the shard is the whole
operator, or a slab whose halo is empty) against the unsharded kernels and the oracle.  The
multi-rank path is covered by tests/test_gpu_multi.py (NCCL) and tests/test_rowshard_gloo.py."""

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
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_whole_operator_as_one_shard_matches_unsharded(dtype, reortho):
    from matfree_b200 import workloads

    m = mfb()
    shape = (9, 8, 7)
    n = int(np.prod(shape))
    k = 14
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, dtype=np.dtype(dtype).name)
    ips, ixs, ds = workloads.laplacian_csr_rows(shape, 0, n, shift=1.0, dtype=np.dtype(dtype).name)
    assert torch.equal(ip, ips) and torch.equal(ix.long(), ixs) and torch.equal(d, ds)
    op = m.ops.csr(ip, ix, d)
    sop = m.ops.csr_row_sharded(ips, ixs, ds, n, 0)
    v = oprng.normal(oprng.prng_key(3), (n,), dtype)
    tri = m.decomp.tridiag_sym(k, reortho=reortho, materialize=False)
    Q1, (d1, e1), r1, c1 = tri(op, v)
    Q2, (d2, e2), r2, c2 = tri(sop, v)
    tol = 2e-5 if dtype == np.float32 else 1e-11
    assert np.allclose(d2.cpu(), d1.cpu(), rtol=tol, atol=tol)
    assert np.allclose(e2.cpu(), e1.cpu(), rtol=tol, atol=tol)
    assert np.allclose(float(c2), float(c1), rtol=1e-6)
    if reortho == "full":
        assert np.allclose(Q2.cpu(), Q1.cpu(), atol=50 * tol)
    # and the oracle
    import scipy.sparse as sp

    A = sp.csr_matrix((d.numpy(), ix.numpy(), ip.numpy()), shape=(n, n))
    od, oe = (ref.lanczos_full_batched if reortho == "full" else ref.lanczos_none_batched)(
        lambda X: (A @ X.T).T, v[None, :], k)[:2]
    assert np.allclose(d2.cpu().numpy(), od[0], rtol=tol, atol=tol)
    assert np.allclose(e2.cpu().numpy(), oe[0][: k - 1], rtol=tol, atol=tol)


def test_sharded_matvec_callable_and_blocks():
    from matfree_b200 import _rowshard, workloads

    m = mfb()
    shape = (6, 5, 4)
    n = int(np.prod(shape))
    ips, ixs, ds = workloads.laplacian_csr_rows(shape, 0, n, shift=0.5)
    sop = m.ops.csr_row_sharded(ips, ixs, ds, n, 0)
    assert sop.shape == (n, n) and sop.plan.halo_rows == 0
    ip, ix, d = workloads.laplacian_csr(shape, shift=0.5)
    op = m.ops.csr(ip, ix, d)
    v = oprng.normal(oprng.prng_key(5), (n,), np.float32)
    assert np.allclose(sop(v).cpu(), op(v).cpu(), rtol=1e-6, atol=1e-6)
    # blocked drivers with ld = 8 start vectors
    V = torch.as_tensor(oprng.normal(oprng.prng_key(6), (n, 8), np.float32)).cuda()
    a, b, ln, Q, res = _rowshard.lanczos_full_sharded(sop, V, 10)
    a1, b1, ln1, Q1, res1 = m.decomp.lanczos_blocked(op, V, 10, "full", want_Q=True, want_residual=True)
    assert np.allclose(a.cpu(), a1.cpu(), rtol=2e-5, atol=2e-5)
    assert np.allclose(b.cpu(), b1.cpu(), rtol=2e-5, atol=2e-5)
    assert np.allclose(res.cpu(), res1.cpu(), atol=1e-4)


def test_sharded_slq_integrand_and_funm_action():
    """The SLQ integrand and `funm_lanczos_sym` accept the row-sharded operator unchanged."""
    from matfree_b200 import workloads

    m = mfb()
    shape = (8, 8, 4)
    n = int(np.prod(shape))
    ips, ixs, ds = workloads.laplacian_csr_rows(shape, 0, n, shift=1.0)
    sop = m.ops.csr_row_sharded(ips, ixs, ds, n, 0)
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0)
    op = m.ops.csr(ip, ix, d)
    v = oprng.rademacher(oprng.prng_key(1), (n,), np.float32)
    integrand = m.funm.integrand_funm_sym_logdet(m.decomp.tridiag_sym(12, reortho="none"))
    assert np.allclose(float(integrand(sop, v)), float(integrand(op, v)), rtol=1e-5)
    act = m.funm.funm_lanczos_sym(m.funm.dense_funm_sym_eigh(("exp", -0.1)), m.decomp.tridiag_sym(12, reortho="full"))
    assert np.allclose(act(sop, v).cpu(), act(op, v).cpu(), rtol=1e-4, atol=1e-5)

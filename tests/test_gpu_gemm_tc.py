"""tcgen05 3xTF32 dense / Gram contraction against fp64 NumPy and the CUDA-core kernel.

The reference's dense matvec is XLA's fp32 `dot_general` (matfree/stochtrace.py:47-49 with
tutorials/1_log_determinants.py:19-21); the bar is fp32-level agreement with the fp64 product.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402


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


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("n,P", [(256, 32), (200, 64), (1000, 128), (520, 256), (1000, 300)])
def test_dense_tcgen05_matches_fp64(variant, n, P):
    from matfree_b200 import _lib

    m = mfb()
    lib = _lib.load()
    A = oprng.normal(oprng.prng_key(7), (n, n), np.float32)
    A = (A + A.T) / 2
    V = oprng.normal(oprng.prng_key(2), (P, n), np.float32)
    want = (A.astype(np.float64) @ V.T.astype(np.float64)).T
    op = m.ops.dense(A)
    assert op._planes is not None
    _lib.check(lib.mf_gemm_config(variant, 1))
    n0 = lib.mf_launch_count()
    got = op.matmat(V).cpu().numpy()
    assert lib.mf_launch_count() > n0
    _lib.check(lib.mf_gemm_config(variant, 0))
    simt = op.matmat(V).cpu().numpy()
    e_tc, e_simt = rel_err(got, want), rel_err(simt, want)
    # 3xTF32 carries ~2^-21 per product; plain fp32 accumulation gives e_simt
    assert e_tc < 2e-6, (e_tc, e_simt)
    assert e_tc < 8 * max(e_simt, 1e-7), (e_tc, e_simt)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("mrows,n,P", [(300, 200, 64), (1031, 512, 256), (4096, 96, 32)])
def test_gram_tcgen05_matches_fp64(variant, mrows, n, P):
    from matfree_b200 import _lib

    m = mfb()
    lib = _lib.load()
    B = oprng.normal(oprng.prng_key(4), (mrows, n), np.float32) / np.float32(np.sqrt(mrows))
    V = oprng.normal(oprng.prng_key(2), (P, n), np.float32)
    B64 = B.astype(np.float64)
    want = ((B64.T @ B64) @ V.T.astype(np.float64)).T
    op = m.ops.gram(B)
    _lib.check(lib.mf_gemm_config(variant, 1))
    got = op.matmat(V).cpu().numpy()
    _lib.check(lib.mf_gemm_config(variant, 0))
    simt = op.matmat(V).cpu().numpy()
    e_tc, e_simt = rel_err(got, want), rel_err(simt, want)
    assert e_tc < 4e-6, (e_tc, e_simt)
    assert e_tc < 8 * max(e_simt, 1e-7), (e_tc, e_simt)


def test_dense_single_vector_call_uses_cuda_core_path():
    """ld = 1 is outside the tensor-core shapes; the callable still works (CUDA cores)."""
    m = mfb()
    A = oprng.normal(oprng.prng_key(7), (64, 64), np.float32)
    v = oprng.normal(oprng.prng_key(2), (64,), np.float32)
    got = m.ops.dense(A)(v).cpu().numpy()
    assert np.allclose(got, A @ v, rtol=1e-5, atol=1e-5)

"""Hutchinson integrands that return per-row values: diagonal,
trace_and_diagonal, rownorms_squared, frobeniusnorm_squared -- block route (mf_probe_gen ->
mf_matmat -> mf_hutch_rows / mf_block_dot) against the oracle on the same keys.

Reference: matfree/stochtrace.py:836-914; its tests
tests/test_stochtrace/test_monte_carlo/test_{diagonal,trace_and_diagonal,rownorms_squared,
frobeniusnorm_squared}.py compare against the dense ground truth at rtol 0.05.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import prng as oprng  # noqa: E402
from oracle import ref  # noqa: E402

RTOL = {np.float32: 2e-5, np.float64: 1e-11}


def mfb():
    import matfree_b200

    return matfree_b200


def _operators(dtype):
    import scipy.sparse as sp

    from matfree_b200 import workloads

    m = mfb()
    shape = (12, 11)
    n = shape[0] * shape[1]
    ip, ix, d = workloads.laplacian_csr(shape, shift=0.75, dtype=np.dtype(dtype).name)
    A = sp.csr_matrix((d.numpy(), ix.numpy(), ip.numpy()), shape=(n, n)).toarray()
    yield "csr", m.ops.csr(ip, ix, d), A
    B = oprng.normal(oprng.prng_key(8), (n, n), dtype)
    B = ((B + B.T) / 2).astype(dtype)
    yield "dense", m.ops.dense(B), B


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("sampler_name", ["signs", "normal"])
@pytest.mark.parametrize("P", [7, 64, 300])
def test_row_integrands_match_oracle(dtype, sampler_name, P):
    m = mfb()
    key = m.prng.prng_key(11)
    okey = oprng.prng_key(11)
    rtol = RTOL[dtype] * (50 if sampler_name == "normal" and dtype == np.float32 else 1)
    for name, op, A in _operators(dtype):
        n = A.shape[0]
        sampler = getattr(m.stochtrace, f"sampler_{sampler_name}")(np.ones(n, dtype), num=P)
        osampler = getattr(ref, f"sampler_{sampler_name}")(n, num=P, dtype=dtype)
        mv = lambda v: (A.astype(np.float64) @ v.astype(np.float64)).astype(dtype)  # noqa: E731
        for integ in ("diagonal", "rownorms_squared", "frobeniusnorm_squared", "trace_and_diagonal"):
            est = m.stochtrace.estimator_monte_carlo_mean_and_sem(getattr(m.stochtrace, f"monte_carlo_{integ}")(), sampler)
            oest = ref.estimator_monte_carlo_mean_and_sem(getattr(ref, f"monte_carlo_{integ}")(), osampler)
            mean, sem = est(op, key)
            omean, osem = oest(mv, okey)
            plain = m.stochtrace.estimator_monte_carlo(getattr(m.stochtrace, f"monte_carlo_{integ}")(), sampler)(op, key)
            if isinstance(omean, dict):
                assert set(mean) == set(omean) == {"trace", "diagonal"}
                pairs = [(mean[k], omean[k], sem[k], osem[k], plain[k]) for k in omean]
            else:
                pairs = [(mean, omean, sem, osem, plain)]
            for g, o, gs, os_, pl in pairs:
                g, gs, pl = g.cpu().numpy(), gs.cpu().numpy(), pl.cpu().numpy()
                scale = np.abs(o).max() + 1e-30
                assert g.shape == np.shape(o), (name, integ)
                assert np.max(np.abs(g - o)) <= rtol * scale, (name, integ, np.max(np.abs(g - o)) / scale)
                assert np.array_equal(pl, g), (name, integ)
                # sem: E[x^2] - mean^2 in fp64 against np.std of fp32 samples
                assert np.allclose(gs, os_, rtol=2e-3, atol=2e-3 * (np.abs(os_).max() + 1e-30)), (name, integ)


def test_diagonal_is_exact_for_rademacher_probes_on_a_diagonal_operator():
    """v_i^2 = 1: every sample of v * (D v) equals diag(D) (the reference's exactness argument)."""
    import scipy.sparse as sp

    m = mfb()
    n = 1000
    dvals = np.arange(1, n + 1, dtype=np.float32)
    op = m.ops.csr_from_scipy(sp.diags(dvals).tocsr())
    sampler = m.stochtrace.sampler_signs(np.ones(n, np.float32), num=513)
    est = m.stochtrace.estimator_monte_carlo_mean_and_sem(m.stochtrace.monte_carlo_diagonal(), sampler)
    mean, sem = est(op, m.prng.prng_key(3))
    assert np.array_equal(mean.cpu().numpy(), dvals)
    assert float(sem.abs().max()) == 0.0
    both = m.stochtrace.estimator_monte_carlo(m.stochtrace.monte_carlo_trace_and_diagonal(), sampler)(op, m.prng.prng_key(3))
    assert float(both["trace"]) == float(dvals.sum())


def test_generic_route_for_unregistered_callables_agrees():
    """A plain callable (not a registered operator) takes the sample-by-sample route of the host
    mirror and must give the same estimate as the block route."""
    m = mfb()
    n = 64
    A = oprng.normal(oprng.prng_key(2), (n, n), np.float32)
    op = m.ops.dense((A + A.T) / 2)
    sampler = m.stochtrace.sampler_signs(np.ones(n, np.float32), num=40)
    key = m.prng.prng_key(5)
    for integ in ("diagonal", "trace_and_diagonal", "rownorms_squared", "frobeniusnorm_squared"):
        f = getattr(m.stochtrace, f"monte_carlo_{integ}")()
        fast = m.stochtrace.estimator_monte_carlo(f, sampler)(op, key)
        slow = m.stochtrace.estimator_monte_carlo(f, sampler)(lambda v: op(v), key)
        if isinstance(fast, dict):
            for k in fast:
                assert np.allclose(fast[k].cpu(), slow[k].cpu(), rtol=1e-4, atol=1e-4)
        else:
            assert np.allclose(fast.cpu(), slow.cpu(), rtol=1e-4, atol=1e-4)

"""Multi-GPU consistency checks on a live process group (one process per GPU, NCCL): the
sharded paths against the single-GPU paths of the same process (which the `-m gpu` tests pin to
the oracle).  Used by `tests/test_gpu_multi.py` (asserted) and by `bench.py` at N > 1 (reported
in the JSON line, so that the driver's scaling run carries multi-GPU parity evidence).
Every function is collective and returns a dict of booleans / numbers; nothing raises on a mismatch.
"""

from __future__ import annotations

import os

import numpy as np

import matfree_b200 as m
from matfree_b200 import _rowshard, workloads


def probe_sharding(dev, shape=(96, 96), P=300, k=12):
    """Probe sharding (`stochtrace.probe_sharding`): per-probe values, mean and sem of an SLQ
    log-determinant and a Hutchinson trace, sharded over the group vs all probes on this GPU."""
    import torch

    n = shape[0] * shape[1]
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, device=dev)
    op = m.ops.csr(ip, ix, d)
    key = m.prng.prng_key(7)
    sampler = m.stochtrace.sampler_signs(np.broadcast_to(np.float32(1), (n,)), num=P)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho="none"))
    est = m.stochtrace.estimator_monte_carlo_mean_and_sem(integrand, sampler)
    plain = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    single = plain.per_probe(op, key, tile=64)
    mean1, sem1 = est(op, key)
    with m.stochtrace.probe_sharding():
        sharded = plain.per_probe(op, key, tile=64)
        mean2, sem2 = est(op, key)
    tr = m.stochtrace.estimator_monte_carlo(m.stochtrace.monte_carlo_trace(), sampler)
    t1 = float(tr(op, key))
    with m.stochtrace.probe_sharding():
        t2 = float(tr(op, key))
    return {
        "per_probe_bit_identical": bool(sharded.shape == single.shape == (P,) and torch.equal(sharded, single)),
        "mean_sem_bit_identical": bool(float(mean1) == float(mean2) and float(sem1) == float(sem2)),
        "trace_bit_identical": bool(t1 == t2),
    }


def _all_ranks_equal(t):
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    allv = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allv, t.contiguous())
    return all(torch.equal(x, allv[0]) for x in allv)


def row_sharding(dev, shape=(24, 16, 16), k=20, block=8):
    """Row sharding (`ops.csr_row_sharded`): `tridiag_sym` (both reortho modes) on slabs of rows
    over the group vs the whole operator on this GPU; the peer-memory route (`mf_lanczos_sharded`)
    vs the NCCL route; scalars bit-identical on all ranks and run to run; the row-sharded SLQ
    estimate (slab probe generation) vs the single-GPU estimate."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(np.prod(shape))
    plane = int(np.prod(shape[1:]))
    r0, r1 = _rowshard.slab_range(n, world, rank, align=plane)
    ip, ix, d = workloads.laplacian_csr_rows(shape, r0, r1, shift=1.0, device=dev)
    sop = m.ops.csr_row_sharded(ip, ix, d, n, r0)
    ipf, ixf, df = workloads.laplacian_csr(shape, shift=1.0, device=dev)
    op = m.ops.csr(ipf, ixf, df)
    out = {"halo_plan_ok": bool(sop.plan.lo == (plane if rank > 0 else 0)
                                and sop.plan.hi == (plane if rank < world - 1 else 0))}
    v = m.prng.rademacher(m.prng.prng_key(1), shape=(n,), dtype=np.float32)   # probe 0 of PRNGKey(1)
    ok = True
    for reortho in ("full", "none"):
        tri = m.decomp.tridiag_sym(k, reortho=reortho, materialize=False)
        Q1, (d1, e1), res1, c1 = tri(op, v)
        Q2, (d2, e2), res2, c2 = tri(sop, v[r0:r1].contiguous())
        ok &= bool(torch.allclose(d2, d1, rtol=1e-5, atol=1e-5) and torch.allclose(e2, e1, rtol=1e-5, atol=1e-5))
        ok &= bool(np.isclose(float(c2), float(c1), rtol=1e-6))
        if reortho == "full":
            ok &= bool(torch.allclose(Q2, Q1[:, r0:r1], atol=1e-4) and torch.allclose(res2, res1[r0:r1], atol=1e-3))
    out["tridiag_matches_single_gpu"] = ok
    out["matvec_bit_identical"] = bool(torch.equal(op(v)[r0:r1], sop(v[r0:r1].contiguous())))
    V = m.prng.normal(m.prng.prng_key(9), shape=(n, block), dtype=np.float32)
    Vloc = V[r0:r1].contiguous()
    routes_ok, ranks_ok, rerun_ok = True, True, True
    peer = _rowshard._use_peer_memory(None)
    for reortho in ("full", "none"):
        a1, b1, l1, Q1, res1 = m.decomp.lanczos_blocked(op, V, k, reortho, want_Q=True, want_residual=True)
        a2, b2, l2, Q2, res2 = m.decomp.lanczos_blocked(sop, Vloc, k, reortho, want_Q=True, want_residual=True)
        if getattr(sop, "_comm", None) is not None:
            sop._comm.check()   # raises if an in-kernel wait timed out
        os.environ["MF_ROWSHARD_NCCL"] = "1"
        try:
            a3, b3, l3, Q3, res3 = m.decomp.lanczos_blocked(sop, Vloc, k, reortho, want_Q=True, want_residual=True)
        finally:
            del os.environ["MF_ROWSHARD_NCCL"]
        for x2, x1, x3 in ((a2, a1, a3), (b2, b1, b3), (l2, l1, l3)):
            routes_ok &= bool(torch.allclose(x2, x1, rtol=2e-5, atol=2e-5) and torch.allclose(x2, x3, rtol=2e-5, atol=2e-5))
        routes_ok &= bool(torch.allclose(Q2, Q1[:, r0:r1], atol=2e-4) and torch.allclose(res2, res1[r0:r1], atol=2e-3))
        ranks_ok &= _all_ranks_equal(torch.cat([a2.flatten(), b2.flatten(), l2.flatten()]))
        a4, b4, _, _, _ = m.decomp.lanczos_blocked(sop, Vloc, k, reortho, want_Q=False, want_residual=False)
        rerun_ok &= bool(torch.equal(a4, a2) and torch.equal(b4, b2))
    out["peer_route_active"] = bool(peer)
    out["peer_and_nccl_routes_match_single_gpu"] = routes_ok
    out["bit_identical_ranks"] = ranks_ok
    out["bit_identical_reruns"] = rerun_ok
    # the row-sharded estimator: slab probes of the global counter -> sharded Lanczos -> quadrature
    P = 24
    sampler = m.stochtrace.sampler_signs(np.broadcast_to(np.float32(1), (n,)), num=P)
    integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho="none"))
    plain = m.stochtrace.estimator_monte_carlo(integrand, sampler)
    q1 = plain.per_probe(op, m.prng.prng_key(3), tile=8)
    q2 = plain.per_probe(sop, m.prng.prng_key(3), tile=8)
    tr = m.stochtrace.estimator_monte_carlo(m.stochtrace.monte_carlo_trace(), sampler)
    t1, t2 = tr.per_probe(op, m.prng.prng_key(3)), tr.per_probe(sop, m.prng.prng_key(3))
    out["sharded_estimator_matches_single_gpu"] = bool(
        torch.allclose(q2, q1, rtol=2e-5) and torch.allclose(t2, t1, rtol=1e-5) and _all_ranks_equal(q2))
    return out


def c4_rowshard(dev, grid=256, depth=100, steps=2, warmup=1, hbm_gbs=6650.0):
    """BASELINE config 4 on the group: `tridiag_sym(reortho="full")`, depth 100, 3-D 7-point
    Laplacian 256^3 row-sharded in slabs of planes; peer-memory route and NCCL route.  Times with
    CUDA events (max over ranks); bytes of SURVEY.md section 8(d): sum_i [4(i+1)+9] n s + matrix."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    g = grid
    shape = (g, g, g)
    n, plane, k = g ** 3, g * g, depth
    r0, r1 = _rowshard.slab_range(n, world, rank, align=plane)
    ip, ix, d = workloads.laplacian_csr_rows(shape, r0, r1, shift=1.0, device=dev)
    nnz_local = int(d.numel())
    op = m.ops.csr_row_sharded(ip, ix, d, n, r0)
    del ix
    torch.cuda.empty_cache()
    v = torch.empty((r1 - r0, 1), dtype=torch.float32, device=dev)
    m.stochtrace._gen_tile(v, op, {"kind": 0}, m.prng.prng_key(1), 0, 1)   # my slab of probe 0
    v = v[:, 0].contiguous()
    tri = m.decomp.tridiag_sym(k, reortho="full", materialize=False)
    nloc, s = r1 - r0, 4
    matrix = nnz_local * (s + 4) + 4 * (nloc + 1)
    alg = sum((4 * (i + 1) + 9) * nloc * s + matrix for i in range(k))
    res = {"workload": f"C4: tridiag_sym(reortho=full), depth {k}, 3-D 7-pt Laplacian {g}^3 + 1.0*I, fp32, "
                       f"row-sharded x{world} (slabs of {nloc // plane} planes)",
           "algorithmic_bytes_per_gpu": alg}
    outs = {}
    for route in ("peer", "nccl"):
        if route == "nccl":
            os.environ["MF_ROWSHARD_NCCL"] = "1"
        try:
            for _ in range(warmup):
                out = tri(op, v)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = tri(op, v)
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.environ.pop("MF_ROWSHARD_NCCL", None)
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if getattr(op, "_comm", None) is not None:
            op._comm.check()
        _, (diag, off), _, c = out
        outs[route] = torch.cat([diag, off])
        res[route] = {"ms": ms, "achieved_gbs_per_gpu": alg / (ms * 1e-3) / 1e9,
                      "frac_of_hbm": alg / (ms * 1e-3) / 1e9 / hbm_gbs,
                      "bit_identical_ranks": _all_ranks_equal(outs[route])}
    theta = np.linalg.eigvalsh(np.diag(diag.double().cpu().numpy()) + np.diag(off.double().cpu().numpy(), 1)
                               + np.diag(off.double().cpu().numpy(), -1))
    lam = 2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1))
    lo, hi = 3 * lam.min() + 1.0, 3 * lam.max() + 1.0
    res["route"] = "peer" if _rowshard._use_peer_memory(None) else "nccl"
    res["ms"] = res[res["route"]]["ms"]
    res["frac_of_hbm"] = res[res["route"]]["frac_of_hbm"]
    res["bit_identical_ranks"] = bool(res["peer"]["bit_identical_ranks"] and res["nccl"]["bit_identical_ranks"])
    res["parity_ok"] = bool(torch.allclose(outs["peer"], outs["nccl"], rtol=2e-5, atol=2e-5)
                            and theta.min() >= lo - 1e-4 and theta.max() <= hi + 1e-4
                            and abs(float(c) - float(np.sqrt(n))) <= 1e-3)
    res["parity_what"] = ("peer-memory route == NCCL route (2e-5), Ritz values inside the closed-form spectrum, "
                          "|v| = sqrt(n); small-size parity vs the single-GPU path: multi_gpu_checks")
    if getattr(op, "_comm", None) is not None:
        op._comm.close()
    return res

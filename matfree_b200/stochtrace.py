"""Stochastic trace estimation on the GPU -- mirrors the Monte-Carlo part of
`matfree/stochtrace.py`.

* `sampler_signs(*args_like, num=)`, `sampler_normal(*args_like, num=)`
  (`stochtrace.py:927-937,957-977`): the returned ``sample(key)`` gives the
  ``(num, n)`` device array `jax.random.rademacher` / `normal` would produce for
  that key (bit-exact Rademacher; Threefry-2x32, partitionable counters).
* `monte_carlo_trace()` (`stochtrace.py:853-865`), `monte_carlo_diagonal()`,
  `monte_carlo_trace_and_diagonal()`, `monte_carlo_rownorms_squared()`,
  `monte_carlo_frobeniusnorm_squared()` (`stochtrace.py:836-914`).
* `estimator_monte_carlo(integrand, sampler)` and `_mean_and_sem`
  (`stochtrace.py:7-52,55-89`).

When the integrand is one of this package's (SLQ or Hutchinson trace), the
sampler is one of this package's and the matvec is a registered operator,
``estimate(matvec, key)`` runs the fused `mf_estimate` kernel chain: probes are
generated tile by tile directly in the blocked layout and never materialised as
a ``(num, n)`` array -- which is what makes 8192 probes of a 16.7M-row operator
possible at all (the reference would need 550 GB for the samples alone).

Multi-GPU: inside `probe_sharding(group)` every rank evaluates its contiguous
slice of the probe range (the counter-based PRNG makes the slices of the
single-device sample array bit-identical), the per-probe values are
all-gathered over NCCL and every rank performs the same reduction.
"""

from __future__ import annotations

import contextlib
import ctypes

import numpy as np

from matfree_b200 import _device, _lib, _sharding, config as _config, funm as _funm, ops

_PROBE_GROUP = {"group": None, "enabled": False}


@contextlib.contextmanager
def probe_sharding(group=None):
    """Shard the probes of every fused `estimate` call across `group` (default: WORLD)."""
    old = dict(_PROBE_GROUP)
    _PROBE_GROUP.update(group=group, enabled=True)
    try:
        yield
    finally:
        _PROBE_GROUP.update(old)


def _flat_like(*args_like):
    """Length and dtype of the flattened `args_like` (a flat array, or a simple pytree)."""
    leaves = []

    def visit(x):
        if isinstance(x, dict):
            for key in sorted(x):
                visit(x[key])
        elif isinstance(x, (list, tuple)):
            for y in x:
                visit(y)
        else:
            leaves.append(x)

    visit(args_like[0] if len(args_like) == 1 else list(args_like))
    n = 0
    is64 = False
    for leaf in leaves:
        shape = tuple(getattr(leaf, "shape", np.shape(leaf)))
        n += int(np.prod(shape)) if shape else 1
        dt = str(getattr(leaf, "dtype", np.asarray(leaf).dtype))
        is64 = is64 or dt.endswith("float64")
    import torch

    return n, (torch.float64 if is64 else torch.float32)


def _row_sharded(matvec):
    from matfree_b200 import _rowshard

    return isinstance(matvec, _rowshard.RowShardedCsr)


def _gen_tile(V, op, sspec, key, t0, npb):
    """Probes ``t0 .. t0+npb-1`` of the sampler's ``(num, n)`` array into the blocked tile
    ``V[rows][ld]``: all rows for an ordinary operator, this rank's slab of rows for a row-sharded
    one (`mf_probe_gen_rows`: counter ``p * n_global + row`` -- the same stream, so the estimate
    does not depend on how the rows are partitioned)."""
    lib = _lib.load()
    rows, ld = V.shape
    mfdt = _device.mf_dtype(V.dtype)
    flags = _config.prng_flags(V.dtype)
    if _row_sharded(op):
        _lib.check(lib.mf_probe_gen_rows(V.data_ptr(), mfdt, op.n_global, op.r0, rows, ld, t0, npb,
                                         int(key[0]), int(key[1]), sspec["kind"], flags,
                                         _device.stream()))
    else:
        _lib.check(lib.mf_probe_gen(V.data_ptr(), mfdt, _lib.MF_LAYOUT_BLOCKED, rows, ld, t0, npb,
                                    int(key[0]), int(key[1]), sspec["kind"], flags, None,
                                    _device.stream()))


def _sharded_values(ispec, sspec, op, key, *, tile=None, return_coeffs=False):
    """Per-probe SLQ / Hutchinson-trace values on a ROW-SHARDED operator (rows and vectors
    partitioned over the ranks of ``op.group``; every rank evaluates every probe on its slab):
    slab probe generation -> `decomp.lanczos_blocked` (the sharded driver: halo exchange, sums
    all-reduced over peer memory or NCCL) -> `funm.quadrature_blocked`.  Collective call."""
    import torch

    from matfree_b200 import _rowshard, decomp

    n, P, dt = sspec["n"], sspec["num"], op.dtype
    if n != op.n_global:
        raise ValueError(f"sampler draws vectors of length {n}, operator dimension is {op.n_global}")
    if sspec["dtype"] != dt:
        raise TypeError(f"sampler dtype {sspec['dtype']} does not match operator dtype {dt}")
    dev = _device.device()
    slq = ispec["kind"] != "trace"
    if slq:
        k = ispec["num_matvecs"]
        if k < 0 or k > n:
            raise ValueError(decomp._error_num_matvecs(k, maxval=n, minval=0))
        if k < 1:
            raise ValueError("the SLQ integrand needs num_matvecs >= 1")
    ld = int(tile) if tile else _device.ld_for(max(P, 1))
    if slq and ispec["reortho"] == "full":
        ld = _cap_tile_for_basis(ld, ispec["num_matvecs"], op.plan.rows_alloc, dt)
    V = torch.empty((op.n, ld), dtype=dt, device=dev)
    parts, coeffs = [], []
    be = _rowshard.CudaBackend(ld)
    for t0 in range(0, P, ld):
        npb = min(ld, P - t0)
        _gen_tile(V, op, sspec, key, t0, npb)
        if npb < ld:
            V[:, npb:] = 1.0  # padding columns: any non-zero vector keeps the recurrence finite
        if not slq:
            W = op.matmat_blocked(V)
            sums = torch.empty((ld,), dtype=torch.float64, device=dev)
            be.block_dot(V, W, sums)
            _rowshard._all_reduce(sums, op.group)
            parts.append(sums[:npb].to(dt))
            continue
        alphas, betas, init_len, _, _ = decomp.lanczos_blocked(
            op, V, ispec["num_matvecs"], ispec["reortho"], want_Q=False, want_residual=False)
        parts.append(_funm.quadrature_blocked(alphas, betas, init_len, npb, ispec["matfun"]))
        if return_coeffs:
            coeffs.append((alphas.clone(), betas.clone(), init_len.clone()))
    vals = torch.cat(parts) if parts else torch.empty((0,), dtype=dt, device=dev)
    if return_coeffs:
        if not coeffs:
            return vals, None, None, None
        return (vals, torch.stack([c[0] for c in coeffs]), torch.stack([c[1] for c in coeffs]),
                torch.stack([c[2] for c in coeffs]))
    return vals


def _cap_tile_for_basis(ld, k, rows, dtype):
    """Narrow the probe tile until the stored basis ``Q[k][rows][ld]`` of a full
    re-orthogonalisation fits in (60 % of) the free device memory."""
    import torch

    es = torch.empty((), dtype=dtype).element_size()
    free, _ = torch.cuda.mem_get_info()
    while ld > 1 and (k + 4) * rows * ld * es > 0.6 * free:
        ld //= 2
    return ld



def _make_sampler(kind: int, args_like, num: int):
    from matfree_b200.backend import tree

    n, dtype = _flat_like(*args_like)
    num = int(num)
    template = args_like[0] if len(args_like) == 1 else list(args_like)
    # stochtrace.py:957-964: one (num, n) draw, then vmap(unflatten) -- shapes only, no data
    shapes = tree.leaf_shapes(template)
    is_flat = tree.is_leaf(template) and len(shapes[0]) <= 1

    def unflatten_batched(mat):
        if is_flat:
            return mat
        parts, off = [], 0
        for shape in shapes:
            size = int(np.prod(shape)) if shape else 1
            parts.append(mat[:, off:off + size].reshape((mat.shape[0],) + shape))
            off += size
        return tree._rebuild(template, iter(parts))

    def sample(key):
        import torch

        lib = _lib.load()
        out = torch.empty((num, n), dtype=dtype, device=_device.device())
        _lib.check(lib.mf_probe_gen(out.data_ptr(), _device.mf_dtype(dtype),
                                    _lib.MF_LAYOUT_PROBE_MAJOR, n, n, 0, num, int(key[0]),
                                    int(key[1]), kind, _config.prng_flags(dtype), None,
                                    _device.stream()))
        return unflatten_batched(out)

    shape = None  # a single array-like keeps its shape in per-row outputs (ravel_pytree's unflatten)
    if len(args_like) == 1 and not isinstance(args_like[0], (dict, list, tuple)):
        shp = tuple(getattr(args_like[0], "shape", np.shape(args_like[0])))
        shape = shp if len(shp) > 1 else None
    sample._mf_sampler = {"kind": kind, "n": n, "num": num, "dtype": dtype, "shape": shape}
    return sample


def sampler_normal(*args_like, num):
    """Sample from a standard-normal distribution (`stochtrace.py:927-929`)."""
    return _make_sampler(_lib.MF_SAMPLER_NORMAL, args_like, num)


def sampler_signs(*args_like, num):
    """Sample signs uniformly / Rademacher (`stochtrace.py:932-937`; real dtypes only)."""
    return _make_sampler(_lib.MF_SAMPLER_SIGNS, args_like, num)


def monte_carlo_trace():
    """Integrand ``v^T (A v)`` (`stochtrace.py:853-865`)."""

    def integrand(matvec, v, *parameters):
        import torch

        v_flat, Qv_flat, _ = _apply_flat(matvec, v, parameters)
        return torch.dot(v_flat, Qv_flat)

    integrand._mf_integrand = {"kind": "trace"}
    return integrand


def _apply_flat(matvec, v, parameters):
    """``(v_flat, (A v)_flat, unravel)`` for a pytree `v` (`stochtrace.py:844-847,859-863`)."""
    from matfree_b200.backend import tree

    if isinstance(matvec, ops.Operator):
        v_flat, unravel = tree.ravel_pytree(v, matvec.dtype)
        return v_flat, matvec(v_flat, *parameters).reshape(-1), unravel
    v_flat, unravel = tree.ravel_pytree(v)
    Qv_flat, _ = tree.ravel_pytree(matvec(v, *parameters), v_flat.dtype)
    return v_flat, Qv_flat, unravel


def _unflatten_like(flat, like):
    """Inverse of the reference's `ravel_pytree` for the simple pytrees samplers accept."""
    if like is None:
        return flat
    return flat.reshape(tuple(like))


def monte_carlo_diagonal():
    """Integrand ``v * (A v)`` (`stochtrace.py:836-849`): its mean estimates ``diag(A)``."""

    def integrand(matvec, v, *parameters):
        v_flat, Qv_flat, unravel = _apply_flat(matvec, v, parameters)
        return unravel(v_flat * Qv_flat)

    integrand._mf_integrand = {"kind": "diagonal"}
    return integrand


def monte_carlo_trace_and_diagonal():
    """Integrand ``{"trace": v^T A v, "diagonal": v * (A v)}`` (`stochtrace.py:868-883`)."""

    def integrand(matvec, v, *parameters):
        import torch

        v_flat, Qv_flat, unravel = _apply_flat(matvec, v, parameters)
        return {"trace": torch.dot(v_flat, Qv_flat), "diagonal": unravel(v_flat * Qv_flat)}

    integrand._mf_integrand = {"kind": "trace_and_diagonal"}
    return integrand


def monte_carlo_rownorms_squared():
    """Integrand ``(A v)^2`` elementwise (`stochtrace.py:886-898`): its mean estimates the squared
    row norms of ``A``."""

    def integrand(matvec, v, *parameters):
        _, Qv_flat, unravel = _apply_flat(matvec, v, parameters)
        return unravel(Qv_flat * Qv_flat)

    integrand._mf_integrand = {"kind": "rownorms_squared"}
    return integrand


def monte_carlo_frobeniusnorm_squared():
    """Integrand ``|A v|^2`` (`stochtrace.py:901-914`): its mean estimates ``|A|_F^2``."""

    def integrand(matvec, v, *parameters):
        import torch

        _, Qv_flat, _ = _apply_flat(matvec, v, parameters)
        return torch.dot(Qv_flat, Qv_flat)

    integrand._mf_integrand = {"kind": "frobeniusnorm_squared"}
    return integrand


_BLOCK_KINDS = ("diagonal", "trace_and_diagonal", "rownorms_squared", "frobeniusnorm_squared")


def _product_values(integrand, sampler, matvec, key, parameters, *, tile=None):
    """Per-probe values of `funm.monte_carlo_funm_product*` with all probes of a tile advancing
    together: `mf_probe_gen` (blocked) -> Golub-Kahan on the block (`decomp.bidiag_blocked`) ->
    `mf_bidiag_quad`.  None if not applicable."""
    ispec = getattr(integrand, "_mf_integrand", None)
    sspec = getattr(sampler, "_mf_sampler", None)
    if (ispec is None or ispec["kind"] != "product" or sspec is None
            or not isinstance(matvec, ops.RectOperator) or parameters):
        return None
    import torch

    from matfree_b200 import decomp

    lib = _lib.load()
    op = matvec
    n, P, dt = sspec["n"], sspec["num"], op.dtype
    if n != op.n:
        raise ValueError(f"sampler draws vectors of length {n}, operator has {op.n} columns")
    if sspec["dtype"] != dt:
        raise TypeError(f"sampler dtype {sspec['dtype']} does not match operator dtype {dt}")
    k = ispec["num_matvecs"]
    if k > min(op.m, op.n) or k < 0:
        raise ValueError(decomp._error_num_matvecs(k, maxval=min(op.m, op.n), minval=0))
    dev = _device.device()
    mfdt = _device.mf_dtype(dt)
    p0, p1, group, world = 0, P, None, 1
    if _PROBE_GROUP["enabled"]:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            group = _PROBE_GROUP["group"]
            world = dist.get_world_size(group)
            p0, p1 = _sharding.shard_range(P, world, dist.get_rank(group))
    nloc = p1 - p0
    ld = int(tile) if tile else _device.ld_for(max(nloc, 1), cap=64)
    V = torch.empty((n, ld), dtype=dt, device=dev)
    parts = []
    for t0 in range(p0, p1, ld):
        npb = min(ld, p1 - t0)
        _lib.check(lib.mf_probe_gen(V.data_ptr(), mfdt, _lib.MF_LAYOUT_BLOCKED, n, ld, t0, npb,
                                    int(key[0]), int(key[1]), sspec["kind"],
                                    _config.prng_flags(dt), None, _device.stream()))
        if npb < ld:
            V[:, npb:] = 1.0  # padding columns: any non-zero vector keeps the recurrence finite
        alphas, betas, init_len, *_ = decomp.bidiag_blocked(op, V, k, ispec["reortho"])
        parts.append(_funm.product_quadrature_blocked(alphas, betas, init_len, npb, ispec["matfun"]))
    vals = torch.cat(parts) if parts else torch.empty((0,), dtype=dt, device=dev)
    if world > 1:
        vals = _sharding.gather_shards(vals, P, group)
    return vals


def _hutchinson_block(integrand, sampler, matvec, key, parameters, *, tile=None):
    """The Hutchinson integrands that are not a single dot product, probe-blocked: per tile one
    `mf_probe_gen` (blocked layout, bit-identical to the reference's sample array), one block
    product `mf_matmat` and one pass that folds the tile into per-row fp64 sums
    (`mf_hutch_rows`) and/or per-probe column sums (`mf_block_dot`).  Returns
    ``(mean, sem)`` with the integrand's output structure, or None if not applicable."""
    ispec = getattr(integrand, "_mf_integrand", None)
    sspec = getattr(sampler, "_mf_sampler", None)
    if (ispec is None or ispec["kind"] not in _BLOCK_KINDS or sspec is None
            or not isinstance(matvec, ops.Operator) or parameters):
        return None
    import torch

    lib = _lib.load()
    op = matvec
    kind = ispec["kind"]
    n, P, dt = sspec["n"], sspec["num"], op.dtype
    sharded = _row_sharded(op)  # rows partitioned over op.group: every rank sees every probe
    n_total = op.n_global if sharded else op.n
    if n != n_total:
        raise ValueError(f"sampler draws vectors of length {n}, operator dimension is {n_total}")
    n = op.n  # rows held by this rank
    if sspec["dtype"] != dt:
        raise TypeError(f"sampler dtype {sspec['dtype']} does not match operator dtype {dt}")
    dev = _device.device()
    mfdt = _device.mf_dtype(dt)
    p0, p1, group, world = 0, P, None, 1
    if _PROBE_GROUP["enabled"] and not sharded:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            group = _PROBE_GROUP["group"]
            world = dist.get_world_size(group)
            p0, p1 = _sharding.shard_range(P, world, dist.get_rank(group))
    nloc = p1 - p0
    ld = int(tile) if tile else _device.ld_for(max(nloc, 1))
    want_rows = kind != "frobeniusnorm_squared"
    want_cols = kind in ("trace_and_diagonal", "frobeniusnorm_squared")
    V = torch.empty((n, ld), dtype=dt, device=dev)
    rows = torch.zeros((2, n), dtype=torch.float64, device=dev) if want_rows else None
    colvals = []
    bws = _device.workspace(lib.mf_blockvec_workspace_bytes(ld, 4))
    first = True
    for t0 in range(p0, p1, ld):
        npb = min(ld, p1 - t0)
        _gen_tile(V, op, sspec, key, t0, npb)
        W = op.matmat_blocked(V)
        if want_rows:
            A = W if kind == "rownorms_squared" else V
            _lib.check(lib.mf_hutch_rows(A.data_ptr(), W.data_ptr(), mfdt, n, ld, npb, int(not first),
                                         rows[0].data_ptr(), rows[1].data_ptr(), _device.stream()))
        if want_cols:
            A = W if kind == "frobeniusnorm_squared" else V
            sums = torch.empty((ld,), dtype=torch.float64, device=dev)
            _lib.check(lib.mf_block_dot(A.data_ptr(), W.data_ptr(), mfdt, n, ld, sums.data_ptr(),
                                        bws.data_ptr(), bws.numel(), _device.stream()))
            if sharded:  # column sums run over all rows: add the other slabs' parts
                from matfree_b200 import _rowshard

                _rowshard._all_reduce(sums, op.group)
            colvals.append(sums[:npb].to(dt))
        first = False
    out_mean, out_sem = {}, {}
    if want_rows:
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(rows, group=group)
        mean = rows[0] / P
        var = torch.clamp(rows[1] / P - mean * mean, min=0.0)
        like = None if sharded else sspec.get("shape")  # sharded: this rank's slab of rows
        out_mean["rows"] = _unflatten_like(mean.to(dt), like)
        out_sem["rows"] = _unflatten_like((torch.sqrt(var) / np.sqrt(P)).to(dt), like)
    if want_cols:
        vals = torch.cat(colvals) if colvals else torch.empty((0,), dtype=dt, device=dev)
        if world > 1:
            vals = _sharding.gather_shards(vals, P, group)
        out_mean["cols"], out_sem["cols"] = _reduce(vals)
    if kind == "trace_and_diagonal":
        return ({"trace": out_mean["cols"], "diagonal": out_mean["rows"]},
                {"trace": out_sem["cols"], "diagonal": out_sem["rows"]})
    if kind == "frobeniusnorm_squared":
        return out_mean["cols"], out_sem["cols"]
    return out_mean["rows"], out_sem["rows"]


# ----------------------------------------------------------------------------


def _fused_values(integrand, sampler, matvec, key, parameters, *, tile=None, return_coeffs=False):
    """Per-probe integrand values through `mf_estimate`; None if not fusable."""
    ispec = getattr(integrand, "_mf_integrand", None)
    sspec = getattr(sampler, "_mf_sampler", None)
    if ispec is None or sspec is None or not isinstance(matvec, ops.Operator) or parameters:
        return None
    from matfree_b200 import adjoint

    if adjoint._needs_grad(*adjoint.diff_tensors_of(matvec, ())):
        return None  # gradients wrt the operator's values: the differentiable per-sample route
    if ispec["kind"] in _BLOCK_KINDS:
        return None  # handled by _hutchinson_block
    if ispec["kind"] == "product":
        return _product_values(integrand, sampler, matvec, key, parameters, tile=tile)
    if _row_sharded(matvec):
        # the single-GPU kernels of mf_estimate know nothing about halos or slabs
        return _sharded_values(ispec, sspec, matvec, key, tile=tile, return_coeffs=return_coeffs)
    import torch

    lib = _lib.load()
    op = matvec
    if sspec["n"] != op.n:
        raise ValueError(f"sampler draws vectors of length {sspec['n']}, operator dimension is {op.n}")
    if sspec["dtype"] != op.dtype:
        raise TypeError(f"sampler dtype {sspec['dtype']} does not match operator dtype {op.dtype}")
    P = sspec["num"]
    dt = op.dtype
    dev = _device.device()

    # probe range of this rank
    p0, p1 = 0, P
    group = None
    world = 1
    if _PROBE_GROUP["enabled"]:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            group = _PROBE_GROUP["group"]
            world = dist.get_world_size(group)
            p0, p1 = _sharding.shard_range(P, world, dist.get_rank(group))
    nloc = p1 - p0

    if ispec["kind"] == "trace":
        integ, k, rflag, fn, param = _lib.MF_INTEGRAND_TRACE, 0, _lib.MF_REORTHO_NONE, 0, 0.0
        host_fn = None
    else:
        integ = _lib.MF_INTEGRAND_SLQ
        k = ispec["num_matvecs"]
        rflag = _lib.MF_REORTHO_FULL if ispec["reortho"] == "full" else _lib.MF_REORTHO_NONE
        matfun = ispec["matfun"]
        known = _funm._known_fn(matfun) if (not callable(matfun) or _funm._is_hashable(matfun)) else None
        host_fn = None if known is not None else matfun
        fn, param = known if known is not None else (_lib.MF_FN_LOG, 0.0)
        if k < 0 or k > op.n:
            from matfree_b200 import decomp

            raise ValueError(decomp._error_num_matvecs(k, maxval=op.n, minval=0))

    ld = int(tile) if tile else _device.ld_for(max(nloc, 1))
    if not tile and rflag == _lib.MF_REORTHO_FULL and integ == _lib.MF_INTEGRAND_SLQ:
        ld = _cap_tile_for_basis(ld, k, op.n, dt)
    ntiles = max(1, -(-nloc // ld))
    st = op._struct()
    ws_bytes = lib.mf_estimate_workspace_bytes(ctypes.byref(st), ld, k, rflag, integ)
    if ws_bytes < 0:
        _lib.check(-1)
    ws = _device.workspace(ws_bytes)
    quad = torch.empty((max(nloc, 1),), dtype=dt, device=dev)
    need_coeffs = return_coeffs or host_fn is not None
    alphas = betas = lens = None
    if need_coeffs and integ == _lib.MF_INTEGRAND_SLQ:
        alphas = torch.empty((ntiles, k, ld), dtype=dt, device=dev)
        betas = torch.empty((ntiles, k, ld), dtype=dt, device=dev)
        lens = torch.empty((ntiles, ld), dtype=dt, device=dev)
    if nloc > 0:
        _lib.check(lib.mf_estimate(ctypes.byref(st), integ, sspec["kind"],
                                   _config.prng_flags(dt), int(key[0]),
                                   int(key[1]), p0, nloc, ld, k, rflag,
                                   _lib.MF_FN_NONE if host_fn is not None else fn, param,
                                   quad.data_ptr(),
                                   None if alphas is None else alphas.data_ptr(),
                                   None if betas is None else betas.data_ptr(),
                                   None if lens is None else lens.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _device.stream()))
    quad = quad[:nloc]
    if host_fn is not None and nloc > 0:
        # arbitrary matfun: Gauss nodes/weights from the kernel, f applied on the device
        parts = []
        for t in range(ntiles):
            npb = min(ld, nloc - t * ld)
            parts.append(_funm.quadrature_blocked(alphas[t], betas[t], lens[t], npb, host_fn))
        quad = torch.cat(parts)
    if world > 1:
        quad = _sharding.gather_shards(quad, P, group)
    if return_coeffs:
        return quad, alphas, betas, lens
    return quad


def _reduce(values):
    """mean, std(ddof=0)/sqrt(P) through `mf_mc_reduce` (deterministic, fp64 accumulation)."""
    import torch

    lib = _lib.load()
    values = values.contiguous()
    stats = torch.empty((4,), dtype=torch.float64, device=values.device)
    _lib.check(lib.mf_mc_reduce(values.data_ptr(), _device.mf_dtype(values.dtype), values.numel(),
                                stats.data_ptr(), _device.stream()))
    return stats[0].to(values.dtype), stats[2].to(values.dtype)


def _tree_map(fn, pytree):
    from matfree_b200.backend import tree

    return tree.tree_map(fn, pytree)


def _generic_values(integrand, sampler, matvecs, key, parameters):
    """Sample-by-sample evaluation of a user integrand (any callable, pytree samples and
    outputs allowed): the reference's `vmap` (`stochtrace.py:49`) as a loop; stacked along a
    leading sample axis."""
    import torch

    from matfree_b200.backend import tree

    samples = sampler(key)
    leaves = tree.tree_leaves(samples)
    num = int(leaves[0].shape[0]) if leaves else 0
    vals = [integrand(matvecs, tree.tree_map(lambda x: x[p], samples), *parameters) for p in range(num)]
    per_sample = [[_device.as_device(x) for x in tree.tree_leaves(v)] for v in vals]
    stacked = [torch.stack([ps[i] for ps in per_sample]) for i in range(len(per_sample[0]))]
    return tree._rebuild(vals[0], iter(stacked))


def estimator_monte_carlo(integrand, /, sampler):
    """Construct a stochastic trace-/diagonal-estimator (`stochtrace.py:7-52`)."""

    def estimate(matvecs, key, *parameters):
        vals = _fused_values(integrand, sampler, matvecs, key, parameters)
        if vals is not None:
            return _reduce(vals)[0]
        blk = _hutchinson_block(integrand, sampler, matvecs, key, parameters)
        if blk is not None:
            return blk[0]
        Qs = _generic_values(integrand, sampler, matvecs, key, parameters)
        return _tree_map(lambda q: q.mean(dim=0), Qs)

    estimate.per_probe = lambda matvecs, key, *parameters, **kw: _fused_values(
        integrand, sampler, matvecs, key, parameters, **kw)
    return estimate


def estimator_monte_carlo_mean_and_sem(integrand, /, sampler):
    """Estimator returning ``(mean, sem)`` with ``sem = std/sqrt(num)`` (`stochtrace.py:55-89`)."""

    def estimate(matvecs, key, *parameters):
        import torch

        vals = _fused_values(integrand, sampler, matvecs, key, parameters)
        if vals is not None:
            return _reduce(vals)
        blk = _hutchinson_block(integrand, sampler, matvecs, key, parameters)
        if blk is not None:
            return blk
        Qs = _generic_values(integrand, sampler, matvecs, key, parameters)
        return (_tree_map(lambda q: q.mean(dim=0), Qs),
                _tree_map(lambda q: q.std(dim=0, unbiased=False) / np.sqrt(q.shape[0]), Qs))

    return estimate

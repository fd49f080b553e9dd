"""Matrix-free tridiagonalisation on the GPU -- mirrors `matfree/decomp.py`.

`tridiag_sym(num_matvecs, materialize=, reortho=, custom_vjp=)` keeps the
reference's signature and defaults (`matfree/decomp.py:30-37`): the default
``reortho="full"`` is Arnoldi with classical Gram-Schmidt applied twice and
``T = (H + H^T)/2`` (`decomp.py:125-145,426-477`); ``reortho="none"`` is the
three-term Lanczos recurrence (`decomp.py:220-292`).  The returned
``decompose(matvec, vec, *params)`` gives the reference's four-field result
``(Q_tall (k, n), J_small, residual (n,), init_length_inv)``.

`matvec` is a registered operator (`matfree_b200.ops`: the whole recurrence runs in
`mf_lanczos`, `include/matfree_b200.h`) or any callable ``matvec(vec, *params)`` of CUDA tensors
with `vec` any pytree (`decomp.py:156-182`): the callable supplies the product, the library the
rest of the recurrence (`matfree_b200/_generic.py`).
"""

from __future__ import annotations

import ctypes
from typing import NamedTuple

from matfree_b200 import _device, _lib, ops


class _DecompResult(NamedTuple):
    # matfree/decomp.py:15-27
    Q_tall: object
    J_small: object
    residual: object
    init_length_inv: object


def _error_num_matvecs(num, maxval, minval):
    # matfree/decomp.py:753-756
    msg1 = f"Parameter 'num_matvecs'={num} exceeds the acceptable range. "
    msg2 = f"Expected: {minval} <= num_matvecs <= {maxval}."
    return msg1 + msg2


def _todense_tridiag_sym(diag, off_diag):
    # matfree/decomp.py:148-152
    import torch

    k = diag.shape[0]
    T = torch.zeros((k, k), dtype=diag.dtype, device=diag.device)
    idx = torch.arange(k, device=diag.device)
    T[idx, idx] = diag
    if k > 1:
        T[idx[:-1], idx[1:]] = off_diag
        T[idx[1:], idx[:-1]] = off_diag
    return T


def lanczos_blocked(op: ops.Operator, V0b, k: int, reortho: str, *, want_Q: bool,
                    want_residual: bool):
    """Run `mf_lanczos` on a blocked start block ``V0b[n][ld]``.

    Returns ``(alphas [k][ld], betas [k][ld], init_len [ld], Q [k][n][ld] | None,
    residual [n][ld] | None)`` as device tensors.
    """
    import torch

    from matfree_b200 import _generic, _rowshard

    if isinstance(op, _generic.CallableOperator):
        # a user callable supplies the product; the rest of the recurrence is the library's
        if reortho == "full":
            a, b, length, Q, res, _ = _generic.arnoldi(op, V0b.contiguous(), k)
        else:
            a, b, length, Q, res = _generic.lanczos_none(op, V0b.contiguous(), k, want_Q=want_Q)
        return a, b, length, Q, (res if want_residual else None)
    if isinstance(op, _rowshard.RowShardedCsr):
        # rows (and vectors) are partitioned over the ranks of op.group: V0b is this rank's slab
        if reortho == "full":
            return _rowshard.lanczos_full_sharded(op, V0b.contiguous(), k, want_residual=want_residual)
        return _rowshard.lanczos_none_sharded(op, V0b.contiguous(), k, want_Q=want_Q,
                                              want_residual=want_residual)
    lib = _lib.load()
    n, ld = V0b.shape
    dt = V0b.dtype
    dev = V0b.device
    full = reortho == "full"
    rflag = _lib.MF_REORTHO_FULL if full else _lib.MF_REORTHO_NONE
    st = op._struct()
    need_Q = want_Q or full
    ws_bytes = lib.mf_lanczos_workspace_bytes(ctypes.byref(st), ld, k, rflag, int(need_Q))
    if ws_bytes < 0:
        _lib.check(-1)
    ws = _device.workspace(ws_bytes)
    alphas = torch.empty((max(k, 1), ld), dtype=dt, device=dev)
    betas = torch.empty((max(k, 1), ld), dtype=dt, device=dev)
    init_len = torch.empty((ld,), dtype=dt, device=dev)
    Q = torch.empty((k, n, ld), dtype=dt, device=dev) if (need_Q and k > 0) else None
    residual = torch.empty((n, ld), dtype=dt, device=dev) if want_residual else None
    _lib.check(lib.mf_lanczos(ctypes.byref(st), V0b.data_ptr(), ld, k, rflag, alphas.data_ptr(),
                              betas.data_ptr(), init_len.data_ptr(),
                              None if Q is None else Q.data_ptr(),
                              None if residual is None else residual.data_ptr(), ws.data_ptr(),
                              ws.numel(), _device.stream()))
    return alphas[:k], betas[:k], init_len, Q, residual


def tridiag_sym(num_matvecs: int, /, *, materialize: bool = True, reortho: str = "full",
                custom_vjp: bool = True):
    """Construct an implementation of tridiagonalisation (`matfree/decomp.py:30-122`).

    With ``custom_vjp=True`` (default) the result is differentiable through `torch.autograd` with
    the reference's adjoints as the backward pass (`matfree_b200.adjoint`: `_tridiag_adjoint` for
    ``reortho="none"``, `_hessenberg_adjoint` for ``"full"``), with respect to the start vector and
    to the operator's values (dense / CSR) or the callable's parameters.  ``custom_vjp=False``
    means "differentiate the forward pass itself" in the reference; the CUDA kernels have no
    autodiff, so a gradient request then raises.
    """
    if reortho not in ("full", "none"):
        msg = f"reortho={reortho} unsupported. Choose eiter {'full', 'none'}."
        raise ValueError(msg)
    k = int(num_matvecs)

    def decompose(matvec, vec, *params):
        import torch

        from matfree_b200 import _generic

        # decomp.py:156-164: `vec` may be any pytree, `matvec(vec, *params)` any callable of
        # device tensors; a registered operator takes the flat vector and no parameters
        op, vec_t, unravel = _generic.wrap(matvec, vec, params)
        n = vec_t.shape[0]
        n_total = getattr(op, "n_global", n)  # row-sharded operators: vec is this rank's slab
        if k < 0 or k > n_total:
            raise ValueError(_error_num_matvecs(k, maxval=n_total, minval=0))
        if n != op.n:
            raise ValueError(f"vector has length {n}, operator dimension is {op.n}")
        from matfree_b200 import adjoint

        diff = adjoint.diff_tensors_of(op, params)
        if k > 0 and adjoint._needs_grad(vec_t, *diff):
            if not custom_vjp:
                raise NotImplementedError("custom_vjp=False: the CUDA kernels cannot be differentiated by "
                                          "autodiff; use the adjoints (custom_vjp=True)")
            return _tridiag_with_grad(op, vec_t, unravel, params, diff, k, reortho, materialize)
        V0b = vec_t.reshape(n, 1)
        alphas, betas, init_len, Q, residual = lanczos_blocked(
            op, V0b, k, reortho, want_Q=True, want_residual=True)
        diags = alphas[:, 0]
        offdiags = betas[: max(k - 1, 0), 0]
        Q_tall = Q[:, :, 0] if Q is not None else torch.zeros((0, n), dtype=op.dtype, device=vec_t.device)
        if k == 0:
            # decomp.py:233-236 / :438-441: empty factorisation; the reference's "none"
            # variant still returns one Lanczos step's remainder, the "full" one the input.
            if reortho == "none":
                a1, b1, _, _, res1 = lanczos_blocked(op, V0b, 1, "none", want_Q=False, want_residual=True)
                res = res1[:, 0]
            else:
                res = vec_t.clone()
        else:
            res = residual[:, 0]
        matrix = (diags, offdiags)
        if materialize:
            matrix = _todense_tridiag_sym(diags, offdiags)
        inv = 1.0 / init_len[0]
        if reortho == "full":
            # decomp.py:132,142: the full variant is built on `hessenberg`, whose fourth output is
            # already 1/|v|, and returns `init_length_inv=1.0 / norm` -- i.e. the LENGTH.  Kept.
            inv = 1.0 / inv
        return _DecompResult(Q_tall=unravel.batched(Q_tall), J_small=matrix, residual=unravel(res),
                             init_length_inv=inv)

    decompose._mf_spec = {"kind": "tridiag_sym", "num_matvecs": k, "reortho": reortho,
                          "materialize": materialize}
    return decompose


def _grad_spec(op, params, forward, reortho):
    from matfree_b200 import _generic

    spec = {"op": op, "forward": forward, "reortho": reortho, "params": tuple(params)}
    if isinstance(op, _generic.CallableOperator):
        spec["fn_flat_params"], spec["params"] = op.fn_params, op.params
    return spec


def _tridiag_with_grad(op, vec_t, unravel, params, diff, k, reortho, materialize):
    """`tridiag_sym` with the adjoints of `matfree_b200.adjoint` as its autograd backward
    (`decomp.py:179-217` for "none"; `:125-145` on top of the Hessenberg VJP for "full")."""
    import torch

    from matfree_b200 import adjoint

    n = vec_t.shape[0]
    if reortho == "full":
        def forward(v):
            return _hessenberg_forward(op, v, k, second_pass=True)

        Q, H, r, c = adjoint.hessenberg_fn().apply(_grad_spec(op, params, forward, "full"), vec_t, *diff)
        T = 0.5 * (H + H.T)                                       # decomp.py:133
        diags, offdiags = torch.diagonal(T, 0), torch.diagonal(T, 1)
        res, inv = r, 1.0 / c                                      # :142 (the reference's quirk: |v|)
    else:
        def forward(v):
            alphas, betas, _, Qb, residual = lanczos_blocked(op, v.reshape(n, 1).contiguous(), k, "none",
                                                             want_Q=True, want_residual=True)
            x_last = residual[:, 0] / betas[k - 1, 0]               # residual = b_{k-1} v_k (:167)
            return (torch.cat([Qb[:, :, 0], x_last[None]]).contiguous(), alphas[:, 0].clone(),
                    betas[:, 0].clone())

        xs, al, be_ = adjoint.tridiag_fn().apply(_grad_spec(op, params, forward, "none"), vec_t, *diff)
        Q, diags, offdiags = xs[:-1], al, be_[:-1]
        res = be_[-1] * xs[-1]                                     # :167
        inv = 1.0 / torch.linalg.vector_norm(vec_t)                # :176-177
    matrix = (diags, offdiags)
    if materialize:
        matrix = torch.diag(diags) + torch.diag(offdiags, 1) + torch.diag(offdiags, -1)
    return _DecompResult(Q_tall=unravel.batched(Q), J_small=matrix, residual=unravel(res), init_length_inv=inv)


def _hessenberg_forward(op, vec, k, *, second_pass):
    """`_hessenberg_forward` (`decomp.py:426-477`): ``(Q (k, n), H (k, k), residual (n,),
    1/|vec|)`` for a registered operator (`mf_hessenberg`) or a callable (`_generic.arnoldi`)."""
    import torch

    from matfree_b200 import _generic

    n = vec.shape[0]
    dt, dev = op.dtype, vec.device
    if isinstance(op, _generic.CallableOperator):
        _, _, init_len, Q, residual, H = _generic.arnoldi(op, vec.reshape(n, 1).contiguous(), k,
                                                          second_pass=second_pass, want_H=True)
        Qk = Q[:, :, 0] if k > 0 else torch.zeros((0, n), dtype=dt, device=dev)
        return Qk, H[:, :, 0], residual[:, 0], 1.0 / init_len[0]
    lib = _lib.load()
    st = op._struct()
    ws = _device.workspace(lib.mf_hessenberg_workspace_bytes(ctypes.byref(st), 1, k))
    H = torch.zeros((k, k, 1), dtype=dt, device=dev)
    Q = torch.zeros((max(k, 1), n, 1), dtype=dt, device=dev)
    init_len = torch.empty((1,), dtype=dt, device=dev)
    residual = torch.empty((n, 1), dtype=dt, device=dev)
    rflag = _lib.MF_REORTHO_FULL if second_pass else _lib.MF_REORTHO_NONE
    vec = vec.contiguous()
    _lib.check(lib.mf_hessenberg(ctypes.byref(st), vec.data_ptr(), 1, k, rflag, H.data_ptr(),
                                 init_len.data_ptr(), Q.data_ptr(), residual.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _device.stream()))
    return Q[:k, :, 0], H[:, :, 0], residual[:, 0], 1.0 / init_len[0]


def bidiag_blocked(op, V0b, k: int, reortho: str):
    """Golub-Kahan bidiagonalisation of a block of start vectors ``V0b[n][ld]``
    (`matfree/decomp.py:645-735`, the operation order of `step`).

    Returns ``(alphas [k][ld], betas [k][ld], init_len [ld], Us [k][m][ld], Vs [k][n][ld],
    vk [n][ld])``: `alphas` is the diagonal of the upper-bidiagonal ``B``, row ``i < k-1`` of
    `betas` its superdiagonal entry ``B[i][i+1]`` and row ``k-1`` the final ``beta`` (the
    reference keeps these as ``betas[1:]`` and ``beta``); `vk` is the normalised last vector.
    All arithmetic on vectors is the library's: two GEMMs per step (`mf_matmat_rect`), the fused
    update+norm kernel, and the CGS dots / update pair applied twice for ``reortho="full"``."""
    import torch

    from matfree_b200 import _rowshard

    n, ld = V0b.shape
    m = op.m
    dt, dev = V0b.dtype, V0b.device
    be = _rowshard.CudaBackend(ld, max_nq=max(k, 1))
    kk = max(k, 1)
    Us = torch.zeros((kk, m, ld), dtype=dt, device=dev)
    Vs = torch.zeros((kk, n, ld), dtype=dt, device=dev)
    alphas = torch.zeros((kk, ld), dtype=dt, device=dev)
    betas = torch.zeros((kk, ld), dtype=dt, device=dev)
    init_len = torch.empty((ld,), dtype=dt, device=dev)
    one = torch.empty((ld,), dtype=dt, device=dev)
    h = torch.empty((kk, ld), dtype=dt, device=dev)
    sums = torch.zeros((kk, ld), dtype=torch.float64, device=dev)
    sq = torch.zeros((ld,), dtype=torch.float64, device=dev)
    full = reortho == "full"

    def cgs_twice(Q, nq, W):
        """W -= Q^T (Q W), twice (decomp.py:706-709,715-718); the squared norm of the result."""
        for rep in range(2):
            be.reorth_dots(Q, nq, W, sums[:nq])
            be.finalize(sums[:nq], False, value=h[:nq])
            be.reorth_update(Q, nq, h, W, sq if rep == 1 else None)

    be.block_dot(V0b, V0b, sq)
    be.finalize(sq, True, value=init_len)
    vk = torch.empty((n, ld), dtype=dt, device=dev)
    be.scale(V0b, init_len, vk, True)                # decomp.py:660
    be.block_dot(vk, vk, sq)                         # `init` normalises once more (:697)
    be.finalize(sq, True, value=one)
    be.scale(vk, one, Vs[0] if k > 0 else vk, True)
    for i in range(k):
        vi = Vs[i]                                   # :699 (vk was written straight into Vs[i])
        W = op.apply_blocked(vi, trans=False)        # :703  A vk
        if i > 0:
            be.lanczos_update(W, Us[i - 1], betas[i - 1], None, None, W, sq)   # :704
            if full:
                cgs_twice(Us, i, W)                  # :705-709 (rows >= i of Us are zero)
        else:
            be.block_dot(W, W, sq)
        be.finalize(sq, True, value=alphas[i])       # :711
        be.scale(W, alphas[i], Us[i], True)
        Z = op.apply_blocked(Us[i], trans=True)      # :714  A^T uk
        be.lanczos_update(Z, vi, alphas[i], None, None, Z, sq)                 # :715
        if full:
            cgs_twice(Vs, i + 1, Z)                  # :716-720
        be.finalize(sq, True, value=betas[i])        # :722
        if i + 1 < k:
            be.scale(Z, betas[i], Vs[i + 1], True)
        else:
            be.scale(Z, betas[i], vk, True)
    return alphas[:k], betas[:k], init_len, Us[:k], Vs[:k], vk


def _todense_bidiag(d, e):
    import torch

    k = d.shape[0]
    B = torch.zeros((k, k), dtype=d.dtype, device=d.device)
    idx = torch.arange(k, device=d.device)
    B[idx, idx] = d
    if k > 1:
        B[idx[:-1], idx[1:]] = e
    return B


def bidiag(num_matvecs: int, /, materialize: bool = True, reortho: str = "full"):
    """Construct an implementation of bidiagonalisation via the Golub-Kahan algorithm
    (`matfree/decomp.py:608-750`): ``A ~ U B V^T`` for an arbitrary real rectangular matrix.

    The matvec must be a registered rectangular operator (`ops.rect(A)`: the vector-matrix
    product the reference gets from `jax.vjp` needs the operator's buffers)."""
    if reortho not in ("full", "none"):
        raise ValueError(f"reortho={reortho} unsupported. Choose eiter {'full', 'none'}.")
    k = int(num_matvecs)

    def estimate(Av, v0, *parameters):
        import torch

        from matfree_b200 import _generic

        if isinstance(Av, ops.Operator):
            if parameters:
                raise TypeError("registered operators carry their own buffers; extra matvec parameters "
                                "are only supported for callables")
            if not isinstance(Av, ops.RectOperator):
                raise TypeError("bidiag: a registered matvec must be ops.rect(A)")
            op = Av
            v = _device.as_device(v0, op.dtype).reshape(-1)
        elif not callable(Av):
            raise TypeError(f"bidiag: matvec must be ops.rect(A) or a callable, got {type(Av).__name__}")
        else:
            # any callable of device tensors: the vector-matrix product comes from its VJP
            # (decomp.py:703,712 use jax.vjp; here torch.func.vjp)
            v = _device.as_device(v0).reshape(-1)
            op = _generic.CallableRect(Av, v, parameters)
        if v.shape[0] != op.n:
            raise ValueError(f"vector has length {v.shape[0]}, operator has {op.n} columns")
        if k > min(op.m, op.n) or k < 0:
            raise ValueError(_error_num_matvecs(k, maxval=min(op.m, op.n), minval=0))  # decomp.py:655-658
        alphas, betas, init_len, Us, Vs, vk = bidiag_blocked(op, v.reshape(-1, 1).contiguous(), k, reortho)
        d = alphas[:, 0]
        e = betas[: max(k - 1, 0), 0]
        beta = betas[k - 1, 0] if k > 0 else torch.zeros((), dtype=op.dtype, device=v.device)
        J = _todense_bidiag(d, e) if materialize else (d, e)
        return _DecompResult(Q_tall=(Us[:, :, 0], Vs[:, :, 0]), J_small=J, residual=beta * vk[:, 0],
                             init_length_inv=1.0 / init_len[0])

    estimate._mf_spec = {"kind": "bidiag", "num_matvecs": k, "reortho": reortho, "materialize": materialize}
    return estimate


def hessenberg(num_matvecs, /, *, reortho: str, custom_vjp: bool = True, reortho_vjp: str = "match"):
    """Construct a Hessenberg factorisation via the Arnoldi iteration (`matfree/decomp.py:351-477`):
    ``A Q^T ~ Q^T H`` for an arbitrary square operator.

    As in the reference the forward pass runs with ``reortho=reortho_vjp`` (`decomp.py:393-396`):
    with the default "match" the second Gram-Schmidt pass is applied whatever `reortho` says
    (`:466` only tests ``!= "none"``); `reortho` itself selects the adjoint's re-projection
    (`matfree_b200.adjoint`), which is the backward pass of the result under `torch.autograd`
    when ``custom_vjp=True``."""
    if reortho not in ("none", "full"):
        raise TypeError(f"Unexpected input for {reortho}: either of {['none', 'full']} expected.")  # :375-378
    k = int(num_matvecs)

    def estimate(matvec, v, *params):
        import torch

        from matfree_b200 import _generic, _rowshard

        op, vec, unravel = _generic.wrap(matvec, v, params)  # decomp.py:380-386
        if isinstance(op, _rowshard.RowShardedCsr):
            raise NotImplementedError(
                "hessenberg: row-sharded operators are supported by tridiag_sym (both reortho modes), "
                "not by the public Arnoldi factorisation")
        n = vec.shape[0]
        if k < 0 or k > n:
            raise ValueError(_error_num_matvecs(k, maxval=n, minval=0))
        if n != op.n:
            raise ValueError(f"vector has length {n}, operator dimension is {op.n}")
        second = reortho_vjp != "none"
        from matfree_b200 import adjoint

        diff = adjoint.diff_tensors_of(op, params)
        if adjoint._needs_grad(vec, *diff):
            if not custom_vjp:
                raise NotImplementedError("custom_vjp=False: the CUDA kernels cannot be differentiated by "
                                          "autodiff; use the adjoint (custom_vjp=True)")

            def forward(v):
                return _hessenberg_forward(op, v, k, second_pass=second)

            # k = 0 raises in the backward pass, like the reference (decomp.py:483-486)
            Q, H, r, c = adjoint.hessenberg_fn().apply(_grad_spec(op, params, forward, reortho), vec, *diff)
        else:
            Q, H, r, c = _hessenberg_forward(op, vec, k, second_pass=second)
        return _DecompResult(Q_tall=unravel.batched(Q), J_small=H, residual=unravel(r), init_length_inv=c)

    estimate._mf_spec = {"kind": "hessenberg", "num_matvecs": k, "reortho": reortho}
    return estimate

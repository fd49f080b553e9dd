"""Matrix-free tridiagonalisation on the GPU -- mirrors `matfree/decomp.py`.

`tridiag_sym(num_matvecs, materialize=, reortho=, custom_vjp=)` keeps the
reference's signature and defaults (`matfree/decomp.py:30-37`): the default
``reortho="full"`` is Arnoldi with classical Gram-Schmidt applied twice and
``T = (H + H^T)/2`` (`decomp.py:125-145,426-477`); ``reortho="none"`` is the
three-term Lanczos recurrence (`decomp.py:220-292`).  The returned
``decompose(matvec, vec, *params)`` gives the reference's four-field result
``(Q_tall (k, n), J_small, residual (n,), init_length_inv)``.

`matvec` must be a registered operator (`matfree_b200.ops`); the arithmetic
runs in `mf_lanczos` (`include/matfree_b200.h`).
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

    from matfree_b200 import _rowshard

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

    `custom_vjp` is accepted for signature compatibility; gradients are out of
    scope of the B200 hot path.
    """
    del custom_vjp
    if reortho not in ("full", "none"):
        msg = f"reortho={reortho} unsupported. Choose eiter {'full', 'none'}."
        raise ValueError(msg)
    k = int(num_matvecs)

    def decompose(matvec, vec, *params):
        import torch

        if params:
            raise TypeError("registered operators carry their own buffers; extra matvec parameters are not supported")
        op = ops.require_operator(matvec, "tridiag_sym")
        vec_t = _device.as_device(vec, op.dtype).reshape(-1)
        n = vec_t.shape[0]
        n_total = getattr(op, "n_global", n)  # row-sharded operators: vec is this rank's slab
        if k < 0 or k > n_total:
            raise ValueError(_error_num_matvecs(k, maxval=n_total, minval=0))
        if n != op.n:
            raise ValueError(f"vector has length {n}, operator dimension is {op.n}")
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
        return _DecompResult(Q_tall=Q_tall, J_small=matrix, residual=res,
                             init_length_inv=1.0 / init_len[0])

    decompose._mf_spec = {"kind": "tridiag_sym", "num_matvecs": k, "reortho": reortho,
                          "materialize": materialize}
    return decompose

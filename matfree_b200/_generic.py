"""Decompositions for USER CALLABLES ``matvec(v, *params)`` (`matfree/stochtrace.py:47-49`,
`matfree/decomp.py:156-182,380-391`): the reference traces any JAX function; here any function of
CUDA tensors is accepted.  The callable supplies the product, every other vector operation of the
recurrence is the CUDA library's (the C-ABI building blocks `mf_reorth_dots`, `mf_reorth_update`,
`mf_lanczos_update`, `mf_block_dot`, `mf_block_scale`, `mf_sums_finalize`), with the operation
order of the reference.  No fused kernel chain on this route (it needs the operator's buffers):
one product per step and probe, which is what `jax.vmap` of an opaque function costs as well.
"""

from __future__ import annotations

from matfree_b200 import _device, ops
from matfree_b200.backend import tree


class CallableOperator(ops.Operator):
    """A user function ``v_flat -> A v_flat`` on flat CUDA tensors, with the operator protocol the
    block drivers use.  `_struct` raises: the fused routes need a registered operator."""

    kind = -1

    def __init__(self, fn_flat, n, dtype, n_out=None):
        self.fn = fn_flat
        self.n = int(n)
        self.n_out = self.n if n_out is None else int(n_out)
        self.dtype = dtype

    def _struct(self):
        raise TypeError(
            "this matvec is a Python callable: the fused CUDA routes (mf_estimate, mf_lanczos) need a "
            "registered operator (matfree_b200.ops.dense / csr / gram)")

    def __call__(self, v, *params):
        if params:
            raise TypeError("parameters are bound when the callable is wrapped")
        return self.fn(v)

    def matmat_blocked(self, X):
        import torch

        n, ld = X.shape
        W = torch.empty((self.n_out, ld), dtype=X.dtype, device=X.device)
        for c in range(ld):
            W[:, c] = _device.as_device(self.fn(X[:, c].contiguous()), X.dtype).reshape(-1)
        return W


def wrap(matvec, vec, params, dtype=None):
    """``(op, v_flat, unravel)`` for ``matvec(vec, *params)`` with `vec` any pytree: a registered
    operator is returned as is (it takes flat vectors and no parameters), a callable is closed
    over `params` and conjugated with ravel / unravel like `decomp.py:156-164`."""
    if isinstance(matvec, ops.Operator):
        if params:
            raise TypeError("registered operators carry their own buffers; extra matvec parameters "
                            "are only supported for callables")
        v_flat, unravel = tree.ravel_pytree(vec, matvec.dtype)
        return matvec, v_flat, unravel
    if not callable(matvec):
        raise TypeError(f"matvec must be a registered operator or a callable, got {type(matvec).__name__}")
    v_flat, unravel = tree.ravel_pytree(vec, dtype)

    def fn_flat_params(x, *p):
        return tree.ravel_pytree(matvec(unravel(x), *p), x.dtype)[0]

    def fn_flat(x):
        return fn_flat_params(x, *params)

    op = CallableOperator(fn_flat, v_flat.shape[0], v_flat.dtype)
    op.fn_params, op.params = fn_flat_params, tuple(params)  # for the adjoints (parameter VJPs)
    return op, v_flat, unravel


def _backend(ld, k):
    from matfree_b200 import _rowshard

    return _rowshard.CudaBackend(ld, max_nq=max(k, 1))


def arnoldi(op, V0, k, *, second_pass=True, want_H=False):
    """`matfree/decomp.py:426-477` on a block ``V0[n][ld]`` with the product supplied by
    ``op.matmat_blocked``.  Returns ``(alphas, betas, init_len, Q [k][n][ld], residual, H)``:
    alphas / betas as `mf_lanczos(reortho=FULL)` (T = (H + H^T)/2), H ``[k][k][ld]`` if wanted."""
    import torch

    n, ld = V0.shape
    be = _backend(ld, k)
    kk = max(k, 1)
    Q = be.empty((kk, n, ld), V0)
    alphas, betas, h = be.empty((kk, ld), V0), be.empty((kk, ld), V0), be.empty((kk, ld), V0)
    init_len = be.empty((ld,), V0)
    sums, sq = be.sums((kk, ld), V0), be.sums((ld,), V0)
    H = torch.zeros((k, k, ld), dtype=V0.dtype, device=V0.device) if want_H else None
    be.block_dot(V0, V0, sq)
    be.finalize(sq, True, value=init_len)
    cur, length = V0, init_len
    V = V0
    for i in range(k):
        be.scale(cur, length, Q[i], True)                       # decomp.py:456-457
        V = op.matmat_blocked(Q[i]).contiguous()                # :460
        be.reorth_dots(Q, i + 1, V, sums[: i + 1])              # :463 (filled columns only)
        be.finalize(sums[: i + 1], False, value=h[: i + 1])
        alphas[i].copy_(h[i])
        if want_H:
            H[: i + 1, i] = h[: i + 1]
        if i > 0:
            be.full_offdiag(betas[i - 1], h[i - 1])             # T = (H + H^T)/2, :133-135
        if second_pass:
            be.reorth_update(Q, i + 1, h, V)                    # :464
            be.reorth_dots(Q, i + 1, V, sums[: i + 1])          # :468
            be.finalize(sums[: i + 1], False, value=h[: i + 1])
        be.reorth_update(Q, i + 1, h, V, sq)                    # :464 / :468, norm fused (:471)
        be.finalize(sq, True, value=betas[i])
        if want_H and i + 1 < k:
            H[i + 1, i] = betas[i]                              # :474
        cur, length = V, betas[i]
    residual = V if k > 0 else V0.clone()
    return alphas[:k], betas[:k], init_len, Q[:k], residual, H


def lanczos_none(op, V0, k, *, want_Q=True):
    """`matfree/decomp.py:220-292` (normalise, product, alpha, update, beta) on a block."""
    n, ld = V0.shape
    be = _backend(ld, k)
    kk = max(k, 1)
    Q = be.empty((k, n, ld), V0) if (want_Q and k > 0) else None
    alphas, betas = be.empty((kk, ld), V0), be.empty((kk, ld), V0)
    init_len = be.empty((ld,), V0)
    sq = be.sums((ld,), V0)
    bufs = [be.empty((n, ld), V0), be.empty((n, ld), V0)]
    R = be.empty((n, ld), V0)
    be.block_dot(V0, V0, sq)
    be.finalize(sq, True, value=init_len)
    cur, length, prev = V0, init_len, None
    for j in range(k):
        vj = Q[j] if Q is not None else bufs[j % 2]
        be.scale(cur, length, vj, True)                         # v_j = r / b   (:227,291)
        W = op.matmat_blocked(vj).contiguous()                  # :287
        be.block_dot(vj, W, sq)                                 # :288
        be.finalize(sq, False, value=alphas[j])
        be.lanczos_update(W, vj, alphas[j], prev, betas[j - 1] if j > 0 else None, R, sq)  # :289
        be.finalize(sq, True, value=betas[j])                   # :290
        cur, length, prev = R, betas[j], vj
    residual = R if k > 0 else V0.clone()
    return alphas[:k], betas[:k], init_len, Q, residual


class CallableRect:
    """A user function ``v -> A v`` (``n -> m``, flat CUDA tensors after `params` are bound) with
    the two products Golub-Kahan needs: the function itself and its vector-Jacobian product,
    ``u -> A^T u`` (`matfree/decomp.py:703,712` take it from `jax.vjp`; here `torch.func.vjp`)."""

    def __init__(self, fn, v, params):
        import torch

        self._fn = lambda x: fn(x, *params)
        self.n = int(v.shape[0])
        out, self._vjp = torch.func.vjp(self._fn, v)
        self.m = int(out.reshape(-1).shape[0])
        self.dtype = v.dtype

    def apply_blocked(self, X, *, trans: bool):
        import torch

        rows_out = self.n if trans else self.m
        W = torch.empty((rows_out, X.shape[1]), dtype=X.dtype, device=X.device)
        for c in range(X.shape[1]):
            x = X[:, c].contiguous()
            W[:, c] = (self._vjp(x)[0] if trans else self._fn(x)).reshape(-1)
        return W

"""Gradients of the decompositions: the adjoints of the Lanczos and Arnoldi iterations.

Mirrors the custom VJPs of `matfree/decomp.py` -- `_tridiag_adjoint` (`:184-217,295-348`,
`tridiag_sym(reortho="none", custom_vjp=True)`) and `_hessenberg_adjoint` (`:398-423,480-600`,
`hessenberg(custom_vjp=True)` and through it `tridiag_sym(reortho="full")`), Kraemer et al. (2024),
"Gradients of functions of large matrices".  The backward pass is a reverse recurrence of `k`
more operator products plus the same kind of vector work as the forward pass; all of it runs in
the CUDA library: dots `mf_block_dot`, masked projections `mf_reorth_dots` / `mf_reorth_update`,
combinations of the stored bases `mf_basis_combine`, the step formulas `mf_lincomb`, products
`mf_matmat` / `mf_matmat_rect`, and the parameter gradient of a CSR operator `mf_sddmm_csr`.
Only `k`-sized scalar bookkeeping (mu, nu, the `k x k` matrices Gamma, Pi_gamma) is torch's.

The reference gets parameter gradients from `jax.vjp` of the user matvec.  Here:

* registered operators are differentiated with respect to their values -- `ops.dense(A)`: ``dA =
  sum_steps cot arg^T`` (one `k`-deep GEMM at the end); `ops.csr(...)`: the same outer products on
  the sparsity pattern (`mf_sddmm_csr`);
* callables ``matvec(v, *params)`` of CUDA tensors with `torch.func.vjp`, as the reference does
  with `jax.vjp` (`decomp.py:345-346,588-590`).

`decomp.tridiag_sym` / `decomp.hessenberg` route through `TridiagFn` / `HessenbergFn`
(`torch.autograd.Function`) whenever an input requires a gradient and ``custom_vjp=True``, so
``loss.backward()`` through an SLQ estimate works like `jax.grad` does in the reference
(tutorials/8_gaussian_logpdf.py:65-66).
"""

from __future__ import annotations

import ctypes

from matfree_b200 import _device, _lib, ops


# ----------------------------------------------------------------------------- small helpers


class _Vec:
    """The block-vector kernels on single vectors (`ld = 1`)."""

    def __init__(self, n, dtype, device, max_nq=4):
        import torch

        from matfree_b200 import _rowshard

        self.n, self.dtype, self.device = n, dtype, device
        self.be = _rowshard.CudaBackend(1, max_nq=max(max_nq, 1))
        self.lib = _lib.load()
        self.mfdt = _device.mf_dtype(dtype)
        self._sum1 = torch.zeros((1,), dtype=torch.float64, device=device)

    def empty(self, *shape):
        import torch

        return torch.empty(shape, dtype=self.dtype, device=self.device)

    def dot(self, x, y):
        """``x . y`` as a 1-element tensor of the working dtype (fp64 accumulation)."""
        self.be.block_dot(x.view(self.n, 1), y.view(self.n, 1), self._sum1)
        return self._sum1.to(self.dtype, copy=True)

    def dots(self, Q, nq, v):
        """``Q[:nq] @ v`` (`mf_reorth_dots`) as a tensor ``(nq,)`` of the working dtype."""
        import torch

        sums = torch.zeros((max(nq, 1), 1), dtype=torch.float64, device=self.device)
        if nq > 0:
            self.be.reorth_dots(Q.view(-1, self.n, 1), nq, v.view(self.n, 1), sums[:nq])
        return sums[:nq, 0].to(self.dtype)

    def project_out(self, Q, nq, coeffs, v):
        """``v -= sum_j coeffs[j] Q[j]`` in place (`mf_reorth_update`)."""
        if nq > 0:
            self.be.reorth_update(Q.view(-1, self.n, 1), nq, coeffs.reshape(nq, 1).contiguous(),
                                  v.view(self.n, 1))

    def combine(self, Q, coeffs, out=None):
        """``sum_j coeffs[j] Q[j]`` (`mf_basis_combine`)."""
        k = Q.shape[0]
        out = self.empty(self.n) if out is None else out
        if k == 0:
            return out.zero_()
        c = coeffs.to(self.dtype).reshape(k, 1).contiguous()
        _lib.check(self.lib.mf_basis_combine(Q.data_ptr(), c.data_ptr(), None, self.mfdt, self.n, 1, k,
                                             out.data_ptr(), _device.stream()))
        return out

    def lincomb(self, vectors, coeffs, scales, out=None):
        """``sum_t scales[t] * coeffs[t] * vectors[t]`` left to right (`mf_lincomb`); `coeffs[t]`
        is a 1-element device tensor of the working dtype or None."""
        T = len(vectors)
        out = self.empty(self.n) if out is None else out
        keep = [None if c is None else c.to(self.dtype).reshape(1).contiguous() for c in coeffs]
        vp = (ctypes.c_void_p * T)(*[v.data_ptr() for v in vectors])
        cp = (ctypes.c_void_p * T)(*[None if c is None else c.data_ptr() for c in keep])
        hs = (ctypes.c_double * T)(*[float(s) for s in scales])
        _lib.check(self.lib.mf_lincomb(vp, cp, hs, T, out.data_ptr(), self.mfdt, self.n, 1, _device.stream()))
        return out

    def scale_(self, x, s, divide):
        self.be.scale(x.view(self.n, 1), s.to(self.dtype).reshape(1).contiguous(), x.view(self.n, 1), divide)
        return x


# ----------------------------------------------------------------------------- differentiable operators


class _DenseDiff:
    """``matvec(v, A) = A @ v``: products through the library GEMMs, ``dA = sum cot arg^T``."""

    def __init__(self, op):
        self.op = op
        self.cots, self.args = [], []
        self.rect = ops.RectOperator.__new__(ops.RectOperator)
        self.rect.__dict__.update(A=op.A, dtype=op.dtype, m=op.n, n=op.n, _planes=getattr(op, "_planes", None))

    def apply(self, x):
        return self.op.matmat_blocked(x.view(-1, 1)).view(-1)

    def apply_T(self, x):
        return self.rect.apply_blocked(x.view(-1, 1).contiguous(), trans=True).view(-1)

    def accumulate(self, cot, arg):
        self.cots.append(cot.clone())
        self.args.append(arg.clone())

    def finish(self):
        """``dA[n][n] = C^T G`` with ``C, G (k, n)`` the stacked cotangents / arguments: column
        blocks of at most 256 through `mf_matmat_rect` (``W = C^T X`` for ``X = G[:, block]``)."""
        import torch

        A = self.op.A
        n, k = self.op.n, len(self.cots)
        dA = torch.zeros_like(A)
        if k == 0:
            return (dA,)
        C = torch.stack(self.cots)            # (k, n): the rect operator's matrix, m = k rows
        G = torch.stack(self.args)
        crect = ops.RectOperator.__new__(ops.RectOperator)
        crect.__dict__.update(A=C.contiguous(), dtype=self.op.dtype, m=k, n=n, _planes=None)
        ld = _device.ld_for(n)
        for j0 in range(0, n, ld):
            w = min(ld, n - j0)
            X = torch.zeros((k, ld), dtype=A.dtype, device=A.device)
            X[:, :w] = G[:, j0:j0 + w]
            W = crect.apply_blocked(X, trans=True)      # [n][ld] = C^T X
            dA[:, j0:j0 + w] = W[:, :w]
        return (dA,)


class _CsrDiff:
    """``matvec(v, data) = CSR(indptr, indices, data) @ v``: ``d data`` = the outer products on
    the sparsity pattern (`mf_sddmm_csr`)."""

    def __init__(self, op):
        self.op = op
        self.cots, self.args = [], []
        self._t = None

    def apply(self, x):
        return self.op.matmat_blocked(x.view(-1, 1)).view(-1)

    def _transposed(self):
        if self._t is None:
            import torch

            op = self.op
            n = op.n
            counts = (op.indptr[1:] - op.indptr[:-1]).long()
            rows = torch.repeat_interleave(torch.arange(n, device=op.data.device), counts)
            cols = op.indices.long()
            order = torch.argsort(cols * n + rows)
            t_indptr = torch.zeros(n + 1, dtype=torch.int64, device=op.data.device)
            t_indptr[1:] = torch.cumsum(torch.bincount(cols, minlength=n), dim=0)
            self._t = ops.CsrOperator(t_indptr.to(torch.int32), rows[order].to(torch.int32),
                                      op.data.detach()[order].contiguous(), n)
        return self._t

    def apply_T(self, x):
        return self._transposed().matmat_blocked(x.view(-1, 1)).view(-1)

    def accumulate(self, cot, arg):
        self.cots.append(cot.clone())
        self.args.append(arg.clone())

    def finish(self):
        import torch

        op = self.op
        lib = _lib.load()
        n, k = op.n, len(self.cots)
        ddata = torch.zeros_like(op.data)
        mfdt = _device.mf_dtype(op.dtype)
        for i0 in range(0, k, 256):
            kk = min(256, k - i0)
            ld = _device.ld_for(kk)
            C = torch.stack(self.cots[i0:i0 + kk]).contiguous()   # (kk, n) probe-major
            G = torch.stack(self.args[i0:i0 + kk]).contiguous()
            Cb = torch.zeros((n, ld), dtype=op.dtype, device=ddata.device)
            Gb = torch.zeros((n, ld), dtype=op.dtype, device=ddata.device)
            _lib.check(lib.mf_to_blocked(C.data_ptr(), Cb.data_ptr(), mfdt, n, kk, ld, _device.stream()))
            _lib.check(lib.mf_to_blocked(G.data_ptr(), Gb.data_ptr(), mfdt, n, kk, ld, _device.stream()))
            _lib.check(lib.mf_sddmm_csr(op.indptr.data_ptr(), op.indices.data_ptr(), n, Cb.data_ptr(),
                                        Gb.data_ptr(), ld, kk, int(i0 > 0), ddata.data_ptr(), mfdt,
                                        _device.stream()))
        return (ddata,)


class _CallableDiff:
    """``matvec_flat(v, *params)``: products and parameter VJPs with `torch.func.vjp`, as the
    reference does with `jax.vjp`."""

    def __init__(self, fn_flat, params):
        self.fn = fn_flat
        self.params = tuple(params)
        self.grads = None

    def apply(self, x):
        import torch

        with torch.no_grad():
            return self.fn(x, *self.params).reshape(-1).contiguous()

    def apply_T(self, x):
        raise NotImplementedError  # the Arnoldi adjoint takes vector and parameter VJPs together

    def vjp_both(self, arg, cot):
        """``(A^T cot, d params)`` in one VJP (`decomp.py:588-589`)."""
        import torch

        _, vjp = torch.func.vjp(lambda u, *p: self.fn(u, *p).reshape(-1), arg, *self.params)
        out = vjp(cot)
        self._add(out[1:])
        return out[0].contiguous()

    def accumulate(self, cot, arg):
        import torch

        _, vjp = torch.func.vjp(lambda *p: self.fn(arg, *p).reshape(-1), *self.params)
        self._add(vjp(cot))

    def _add(self, inc):
        from matfree_b200.backend import tree

        if self.grads is None:
            self.grads = [tree.tree_map(lambda g: g.clone(), g) for g in inc]
        else:
            self.grads = [_tree_add(a, b) for a, b in zip(self.grads, inc)]

    def finish(self):
        import torch

        from matfree_b200.backend import tree

        if self.grads is None:
            return tuple(tree.tree_map(torch.zeros_like, p) for p in self.params)
        return tuple(self.grads)


def _tree_add(a, b):
    if isinstance(a, dict):
        return {key: _tree_add(a[key], b[key]) for key in a}
    if isinstance(a, (list, tuple)):
        return type(a)(_tree_add(x, y) for x, y in zip(a, b))
    if a is None:
        return None
    return a + b


def diff_operator(op_or_fn, params=()):
    """The differentiable wrapper of a registered operator or of a flat callable."""
    if isinstance(op_or_fn, ops.DenseOperator):
        return _DenseDiff(op_or_fn)
    if isinstance(op_or_fn, ops.CsrOperator) and not hasattr(op_or_fn, "n_global"):
        return _CsrDiff(op_or_fn)
    if isinstance(op_or_fn, ops.Operator):
        raise NotImplementedError(
            f"gradients with respect to {type(op_or_fn).__name__} are not implemented (dense and CSR "
            "operators and callables are)")
    return _CallableDiff(op_or_fn, params)


# ----------------------------------------------------------------------------- the adjoints


def tridiag_adjoint(dop, *, initvec_norm, alphas, betas, xs, dalphas, dbetas, dxs):
    """`_tridiag_adjoint` (`decomp.py:295-348`): ``xs, dxs (k+1, n)``, ``alphas, dalphas (k,)``,
    ``betas, dbetas (k,)`` (last entries belong to the residual).  Returns ``(grad_initvec (n,),
    grad_params)``."""
    import torch

    k, n = alphas.shape[0], xs.shape[1]
    dt, dev = xs.dtype, xs.device
    V = _Vec(n, dt, dev)
    xs = xs.contiguous()
    dxs = dxs.to(dt).contiguous()
    xi = (-dxs[-1]).contiguous()                      # :316 init_val
    lam_plus = torch.zeros((n,), dtype=dt, device=dev)
    for i in reversed(range(k)):                      # scan(reverse=True), :317-319
        x, xplus = xs[i], xs[i + 1]
        a, b = alphas[i:i + 1], betas[i:i + 1]
        V.scale_(xi, b, True)                         # xi /= b                        (:339)
        mu = dbetas[i:i + 1].to(dt) - V.dot(lam_plus, x) + V.dot(xplus, xi)          # :340
        nu = dalphas[i:i + 1].to(dt) + V.dot(x, xi)                                  # :341
        lam = V.lincomb([xi, xplus, x], [None, mu, nu], [-1.0, 1.0, 1.0])           # :342
        Alam = dop.apply(lam)                         # :345
        dop.accumulate(cot=x, arg=lam)                # :345-346: vjp of p -> matvec(lam, p) at x
        xi = V.lincomb([dxs[i], Alam, lam, lam_plus, xplus], [None, None, a, b, b * nu],
                       [-1.0, -1.0, 1.0, 1.0, -1.0])  # :349
        lam_plus = lam
    lambda_1 = xi                                     # the carry's second slot (:317)
    s = V.dot(lambda_1, xs[0])
    inv = 1.0 / initvec_norm.to(dt).reshape(1)
    grad_initvec = V.lincomb([xs[0], lambda_1], [s * inv, inv], [1.0, -1.0])         # :324
    return grad_initvec, dop.finish()


def hessenberg_adjoint(dop, *, Q, H, r, c, dQ, dH, dr, dc, reortho):
    """`_hessenberg_adjoint` (`decomp.py:480-600`).  ``Q, dQ (k, n)`` with the Krylov vectors as
    ROWS (the layout `estimate` returns; the reference's internal ``(n, k)`` transposed), ``H, dH
    (k, k)``, ``r, dr (n,)``, ``c, dc`` scalars.  Returns ``(dv (n,), grad_params)``."""
    import torch

    k, n = Q.shape
    if k == 0:
        raise ValueError("Custom Hessenberg-adjoints are not implemented for num_matvecs = 0.")  # :483-486
    dt, dev = Q.dtype, Q.device
    V = _Vec(n, dt, dev, max_nq=k)
    Q = Q.contiguous()
    dQ = dQ.to(dt).contiguous()
    dH = dH.to(dt)
    dr = dr.to(dt).contiguous()
    r = r.contiguous()
    eye = torch.eye(k, dtype=dt, device=dev)
    tril = torch.tril(torch.ones((k, k), dtype=dt, device=dev))
    lower_mask = tril - 0.5 * eye                                            # lower(ones), :489-494
    gamma = dH[:, -1] - V.dots(Q, k, dr)                                     # :498
    lambda_k = V.lincomb([dr, V.combine(Q, gamma)], [None, None], [1.0, 1.0])   # :499
    Lambda = torch.zeros_like(Q)
    Gamma = torch.zeros((k, k), dtype=dt, device=dev)
    dQtQ = torch.stack([V.dots(Q, k, dQ[i]) for i in range(k)])              # (dQ^T Q)[i][j]
    e11 = torch.zeros((k, k), dtype=dt, device=dev)
    e11[0, 0] = 1.0
    Pi_gamma = -(dc.to(dt) * c.to(dt)) * e11 + H @ dH.T - dQtQ               # :506
    sub = torch.diagonal(H, -1)
    beta_minuses = torch.cat([torch.ones((1,), dtype=dt, device=dev), sub])
    alphas = torch.diagonal(H)
    beta_pluses = H - torch.diag(alphas) - torch.diag(sub, -1)
    for idx in reversed(range(k)):                                           # scan(reverse=True), :538
        q = Q[idx]
        if reortho == "full":                                                # :576-585
            nq = min(idx + 2, k)                                             # reortho_mask: tril(ones, 1)
            coeff = V.dots(Q, nq, lambda_k) - dH[:nq, idx]
            V.project_out(Q, nq, coeff, lambda_k)
        if isinstance(dop, _CallableDiff):
            vecmat = dop.vjp_both(arg=q, cot=lambda_k)                       # :588-590
        else:
            vecmat = dop.apply_T(lambda_k)
            dop.accumulate(cot=lambda_k, arg=q)
        Gamma[idx, :] = lower_mask[idx] * (Pi_gamma[idx] - V.dots(Q, k, vecmat))   # :593-594
        Lambda[idx].copy_(lambda_k)                                          # :597
        g_row = Gamma[idx, :] + Gamma[:, idx]
        t1 = V.combine(Q, g_row)                                             # (Gamma + Gamma^T)[idx] @ Q^T
        t2 = V.combine(Lambda, beta_pluses[idx])                             # beta_plus @ Lambda^T
        lambda_k = V.lincomb([dQ[idx], r, t1, lambda_k, vecmat, t2],
                             [None, gamma[idx:idx + 1], None, alphas[idx:idx + 1], None, None],
                             [1.0, 1.0, 1.0, -1.0, 1.0, -1.0])               # :598-599
        V.scale_(lambda_k, beta_minuses[idx:idx + 1], True)                  # :600
    dv = V.scale_(lambda_k, c.to(dt).reshape(1), False)                      # :543
    return dv, dop.finish()


# ----------------------------------------------------------------------------- autograd glue


def _needs_grad(*tensors):
    import torch

    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def diff_tensors_of(op, params):
    """The tensors a decomposition is differentiated with respect to besides the start vector:
    a registered operator's values, or the tensor leaves of a callable's parameters."""
    import torch

    from matfree_b200 import _generic
    from matfree_b200.backend import tree

    if isinstance(op, _generic.CallableOperator):
        params = op.params if hasattr(op, "params") else params
        return [leaf for p in params for leaf in tree.tree_leaves(p) if isinstance(leaf, torch.Tensor)]
    if isinstance(op, ops.DenseOperator):
        return [op.A]
    if isinstance(op, ops.CsrOperator) and not hasattr(op, "n_global"):
        return [op.data]
    return []


def _make_dop(spec):
    """(differentiable operator, how to map its gradients to the autograd inputs)."""
    from matfree_b200 import _generic

    op = spec["op"]
    if isinstance(op, _generic.CallableOperator):
        return _CallableDiff(spec["fn_flat_params"], spec["params"])
    return diff_operator(op)


def _param_grads_to_leaves(spec, grads):
    """Gradients in the order of `diff_tensors_of`."""
    from matfree_b200 import _generic
    from matfree_b200.backend import tree

    if isinstance(spec["op"], _generic.CallableOperator):
        import torch

        out = []
        for p, g in zip(spec["params"], grads):
            for leaf, gl in zip(tree.tree_leaves(p), tree.tree_leaves(g)):
                if isinstance(leaf, torch.Tensor):
                    out.append(gl)
        return out
    return list(grads)


def hessenberg_fn():
    import torch

    class HessenbergFn(torch.autograd.Function):
        """`decomp.hessenberg`'s forward pass with `_hessenberg_adjoint` as its backward."""

        @staticmethod
        def forward(ctx, spec, vec, *diff_tensors):
            Q, H, r, c = spec["forward"](vec.detach())
            ctx.spec = spec
            ctx.save_for_backward(Q, H, r, c)
            return Q, H, r, c

        @staticmethod
        def backward(ctx, dQ, dH, dr, dc):
            Q, H, r, c = ctx.saved_tensors
            spec = ctx.spec
            z = torch.zeros_like
            dQ, dH, dr, dc = (z(t) if d is None else d for t, d in zip((Q, H, r, c), (dQ, dH, dr, dc)))
            dv, grads = hessenberg_adjoint(_make_dop(spec), Q=Q, H=H, r=r, c=c, dQ=dQ, dH=dH, dr=dr, dc=dc,
                                           reortho=spec["reortho"])
            return (None, dv, *_param_grads_to_leaves(spec, grads))

    return HessenbergFn


def tridiag_fn():
    import torch

    class TridiagFn(torch.autograd.Function):
        """`_tridiag_forward` (`reortho="none"`) with `_tridiag_adjoint` as its backward; outputs
        ``xs (k+1, n)`` (all Lanczos vectors incl. the last), ``alphas (k,)``, ``betas (k,)``."""

        @staticmethod
        def forward(ctx, spec, vec, *diff_tensors):
            xs, alphas, betas = spec["forward"](vec.detach())
            ctx.spec = spec
            ctx.save_for_backward(xs, alphas, betas, torch.linalg.vector_norm(vec.detach()))
            return xs, alphas, betas

        @staticmethod
        def backward(ctx, dxs, dalphas, dbetas):
            xs, alphas, betas, norm = ctx.saved_tensors
            spec = ctx.spec
            z = torch.zeros_like
            dxs, dalphas, dbetas = (z(t) if d is None else d
                                    for t, d in zip((xs, alphas, betas), (dxs, dalphas, dbetas)))
            gv, grads = tridiag_adjoint(_make_dop(spec), initvec_norm=norm, alphas=alphas, betas=betas, xs=xs,
                                        dalphas=dalphas, dbetas=dbetas, dxs=dxs)
            return (None, gv, *_param_grads_to_leaves(spec, grads))

    return TridiagFn

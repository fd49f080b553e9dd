"""Matrix functions via Lanczos on the GPU -- mirrors the SLQ part of `matfree/funm.py`.

* `monte_carlo_funm_sym_logdet(tridiag_sym)` / `monte_carlo_funm_sym(dense_funm,
  tridiag_sym)` (`funm.py:186-243`) -- integrands for
  `stochtrace.estimator_monte_carlo`; also exported under the names
  `integrand_funm_sym_logdet` / `integrand_funm_sym`.
* `dense_funm_sym_eigh(matfun)` (`funm.py:322-335`).
* `funm_lanczos_sym(dense_funm, tridiag_sym)` (`funm.py:114-147`).

The quadrature ``|v|^2 e1^T f(T) e1`` is computed by `mf_tridiag_quad` (implicit
QL on the tridiagonal, first eigenvector row only).  Recognised `matfun`s (log,
exp, sqrt, sin, reciprocal, identity, powers) are fused into that kernel; any
other elementwise callable is applied on the device to the `(P, k)` array of
Ritz values the kernel returns (Gauss nodes / weights), which is mathematically
the same contraction.
"""

from __future__ import annotations

import ctypes
import math

import numpy as np

from matfree_b200 import _device, _lib, decomp, ops


def _known_fn(matfun):
    """Map a callable / name to (fn_id, param) of the fused kernel, or None."""
    if isinstance(matfun, str):
        table = {"log": (_lib.MF_FN_LOG, 0.0), "exp": (_lib.MF_FN_EXP, 1.0),
                 "sqrt": (_lib.MF_FN_SQRT, 0.0), "inv": (_lib.MF_FN_INV, 0.0),
                 "identity": (_lib.MF_FN_IDENTITY, 0.0), "sin": (_lib.MF_FN_SIN, 1.0)}
        if matfun not in table:
            raise ValueError(f"unknown matrix function {matfun!r}")
        return table[matfun]
    if isinstance(matfun, tuple) and len(matfun) == 2 and isinstance(matfun[0], str):
        name, param = matfun
        ids = {"exp": _lib.MF_FN_EXP, "pow": _lib.MF_FN_POW, "sin": _lib.MF_FN_SIN}
        if name not in ids:
            raise ValueError(f"unknown parametrised matrix function {name!r}")
        return ids[name], float(param)
    known = {np.log: (_lib.MF_FN_LOG, 0.0), math.log: (_lib.MF_FN_LOG, 0.0),
             np.exp: (_lib.MF_FN_EXP, 1.0), math.exp: (_lib.MF_FN_EXP, 1.0),
             np.sqrt: (_lib.MF_FN_SQRT, 0.0), math.sqrt: (_lib.MF_FN_SQRT, 0.0),
             np.sin: (_lib.MF_FN_SIN, 1.0), math.sin: (_lib.MF_FN_SIN, 1.0),
             np.reciprocal: (_lib.MF_FN_INV, 0.0)}
    try:
        import torch

        known.update({torch.log: (_lib.MF_FN_LOG, 0.0), torch.exp: (_lib.MF_FN_EXP, 1.0),
                      torch.sqrt: (_lib.MF_FN_SQRT, 0.0), torch.sin: (_lib.MF_FN_SIN, 1.0),
                      torch.reciprocal: (_lib.MF_FN_INV, 0.0)})
    except Exception:  # pragma: no cover
        pass
    try:
        return known.get(matfun)
    except TypeError:  # unhashable callable
        return None


def dense_funm_sym_eigh(matfun):
    """Dense matrix function via a symmetric eigendecomposition (`funm.py:322-335`).

    The returned callable works on a dense symmetric device matrix.  It also
    records `matfun` so the SLQ integrand can fuse it into the quadrature kernel.
    """

    def fun(dense_matrix):
        import torch

        M = _device.as_device(dense_matrix)
        eigvals, eigvecs = torch.linalg.eigh(M)
        fx = _apply_matfun(matfun, eigvals)
        return eigvecs @ torch.diag(fx) @ eigvecs.T

    fun._mf_matfun = matfun
    return fun


def _apply_matfun(matfun, x):
    """Apply an elementwise `matfun` to a device tensor."""
    import torch

    known = _known_fn(matfun) if not callable(matfun) or _is_hashable(matfun) else None
    if known is not None:
        fn, param = known
        return {
            _lib.MF_FN_LOG: lambda t: torch.log(t),
            _lib.MF_FN_EXP: lambda t: torch.exp(param * t),
            _lib.MF_FN_INV: lambda t: 1.0 / t,
            _lib.MF_FN_SQRT: lambda t: torch.sqrt(t),
            _lib.MF_FN_POW: lambda t: t ** param,
            _lib.MF_FN_IDENTITY: lambda t: t,
            _lib.MF_FN_SIN: lambda t: torch.sin(param * t),
        }[fn](x)
    try:
        out = matfun(x)
        if isinstance(out, torch.Tensor):
            return out
    except Exception:
        pass
    return torch.as_tensor(np.asarray(matfun(x.detach().cpu().numpy())), device=x.device).to(x.dtype)


def _is_hashable(obj):
    try:
        hash(obj)
        return True
    except TypeError:
        return False


def quadrature_blocked(alphas, betas, init_len, num_probes, matfun):
    """``init_len^2 * e1^T f(T) e1`` for every probe of a tile -> tensor ``[num_probes]``."""
    import torch

    lib = _lib.load()
    k, ld = alphas.shape
    dt = alphas.dtype
    dev = alphas.device
    ws = _device.workspace(lib.mf_tridiag_quad_workspace_bytes(ld, k))
    known = _known_fn(matfun) if (not callable(matfun) or _is_hashable(matfun)) else None
    quad = torch.empty((ld,), dtype=dt, device=dev)
    if known is not None:
        fn, param = known
        _lib.check(lib.mf_tridiag_quad(alphas.data_ptr(), betas.data_ptr(), init_len.data_ptr(),
                                       _device.mf_dtype(dt), ld, num_probes, k, fn, param,
                                       quad.data_ptr(), None, None, ws.data_ptr(), ws.numel(),
                                       _device.stream()))
        return quad[:num_probes]
    nodes = torch.empty((k, ld), dtype=torch.float64, device=dev)
    weights = torch.empty((k, ld), dtype=torch.float64, device=dev)
    _lib.check(lib.mf_tridiag_quad(alphas.data_ptr(), betas.data_ptr(), init_len.data_ptr(),
                                   _device.mf_dtype(dt), ld, num_probes, k, _lib.MF_FN_NONE, 0.0,
                                   None, nodes.data_ptr(), weights.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _device.stream()))
    fx = _apply_matfun(matfun, nodes[:, :num_probes].to(dt)).to(torch.float64)
    q = (fx * weights[:, :num_probes]).sum(dim=0) * init_len[:num_probes].to(torch.float64) ** 2
    return q.to(dt)


def ritz_blocked(alphas, betas, num_probes):
    """Gauss nodes (Ritz values, ascending) and weights, fp64 ``[k][num_probes]``."""
    import torch

    lib = _lib.load()
    k, ld = alphas.shape
    dev = alphas.device
    ws = _device.workspace(lib.mf_tridiag_quad_workspace_bytes(ld, k))
    nodes = torch.empty((k, ld), dtype=torch.float64, device=dev)
    weights = torch.empty((k, ld), dtype=torch.float64, device=dev)
    _lib.check(lib.mf_tridiag_quad(alphas.data_ptr(), betas.data_ptr(), None,
                                   _device.mf_dtype(alphas.dtype), ld, num_probes, k,
                                   _lib.MF_FN_NONE, 0.0, None, nodes.data_ptr(), weights.data_ptr(),
                                   ws.data_ptr(), ws.numel(), _device.stream()))
    return nodes[:, :num_probes], weights[:, :num_probes]


def monte_carlo_funm_sym(dense_funm, tridiag_sym, /):
    """Integrand for matrix-function-trace estimation (`funm.py:205-243`)."""
    spec = getattr(tridiag_sym, "_mf_spec", None)
    matfun = getattr(dense_funm, "_mf_matfun", None)

    def quadform(matvec, v0, *parameters):
        if spec is None or matfun is None:
            return _quadform_generic(dense_funm, tridiag_sym, matvec, v0, *parameters)
        from matfree_b200 import _generic

        from matfree_b200 import adjoint

        # funm.py:226-235: any pytree `v0`, any callable `matvec(v0, *parameters)`
        op, v, _ = _generic.wrap(matvec, v0, parameters)
        if adjoint._needs_grad(v, *adjoint.diff_tensors_of(op, parameters)):
            # differentiable route: the decomposition's adjoint (matfree_b200.adjoint) under
            # torch.autograd, the k x k function through torch.linalg.eigh
            return _quadform_generic(dense_funm, tridiag_sym, matvec, v0, *parameters)
        k = spec["num_matvecs"]
        n_total = getattr(op, "n_global", v.shape[0])
        if k < 0 or k > n_total:
            raise ValueError(decomp._error_num_matvecs(k, maxval=n_total, minval=0))
        alphas, betas, init_len, _, _ = decomp.lanczos_blocked(
            op, v.reshape(-1, 1), k, spec["reortho"], want_Q=False, want_residual=False)
        return quadrature_blocked(alphas, betas, init_len, 1, matfun)[0]

    quadform._mf_integrand = None if (spec is None or matfun is None) else {
        "kind": "slq", "num_matvecs": spec["num_matvecs"], "reortho": spec["reortho"],
        "matfun": matfun}
    return quadform


def _quadform_generic(dense_funm, tridiag_sym, matvec, v0, *parameters):
    # funm.py:226-241 with user-supplied pieces (device tensors)
    import torch

    from matfree_b200.backend import tree

    v0_flat, unravel = tree.ravel_pytree(v0)
    length = torch.linalg.vector_norm(v0_flat)

    def matvec_flat(v_f, *p):
        return tree.ravel_pytree(matvec(unravel(v_f), *p))[0]

    mv = matvec if isinstance(matvec, ops.Operator) else matvec_flat
    _, dense, *_ = tridiag_sym(mv, v0_flat / length, *parameters)
    if isinstance(dense, tuple):  # materialize=False
        dense = torch.diag(dense[0]) + torch.diag(dense[1], 1) + torch.diag(dense[1], -1)
    fA = dense_funm(dense)
    return length**2 * fA[0, 0]


def monte_carlo_funm_sym_logdet(tridiag_sym, /):
    """Integrand for the log-determinant (`funm.py:186-202`)."""
    return monte_carlo_funm_sym(dense_funm_sym_eigh(np.log), tridiag_sym)


# BASELINE.json's north_star spells these `integrand_funm_sym[_logdet]`
integrand_funm_sym = monte_carlo_funm_sym
integrand_funm_sym_logdet = monte_carlo_funm_sym_logdet


def funm_lanczos_sym(dense_funm, tridiag_sym, /):
    """Matrix-function-vector product ``f(A) v`` via Lanczos (`funm.py:114-147`).

    With ``reortho="none"`` and a recognised `matfun` the product is formed by the two-pass
    kernel chain `mf_funm_lanczos` (no stored basis: what makes BASELINE config 5 -- 1e7 rows,
    4096 probes -- fit at all); otherwise the basis is stored and contracted
    (`mf_tridiag_funm_e1` + `mf_basis_combine`).  ``estimate.batched(matvec, V)`` applies it to
    every row of ``V (P, n)`` with all probes of a tile advancing together (what `jax.vmap` of
    the reference's function does)."""
    spec = getattr(tridiag_sym, "_mf_spec", None)
    matfun = getattr(dense_funm, "_mf_matfun", None)

    def _known():
        if matfun is not None and (not callable(matfun) or _is_hashable(matfun)):
            return _known_fn(matfun)
        return None

    def _check(op, n):
        k = spec["num_matvecs"]
        n_total = getattr(op, "n_global", n)
        if k < 0 or k > n_total:
            raise ValueError(decomp._error_num_matvecs(k, maxval=n_total, minval=0))
        return k

    def _two_pass_blocked(op, V0b, num_probes, known):
        import torch

        lib = _lib.load()
        n, ld = V0b.shape
        k = spec["num_matvecs"]
        st = op._struct()
        nbytes = lib.mf_funm_lanczos_workspace_bytes(ctypes.byref(st), ld, k)
        if nbytes < 0:
            _lib.check(-1)
        ws = _device.workspace(nbytes)
        out = torch.empty_like(V0b)
        _lib.check(lib.mf_funm_lanczos(ctypes.byref(st), V0b.data_ptr(), ld, num_probes, k, known[0],
                                       known[1], out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _device.stream()))
        return out

    def _use_two_pass(op, known):
        from matfree_b200 import _rowshard

        from matfree_b200 import _generic

        return (known is not None and spec["reortho"] == "none" and spec["num_matvecs"] >= 1
                and not isinstance(op, (_rowshard.RowShardedCsr, _generic.CallableOperator)))

    def estimate(matvec, vec, *parameters):
        import torch

        from matfree_b200 import _generic

        if spec is None:
            raise TypeError("funm_lanczos_sym: tridiag_sym must come from matfree_b200.decomp.tridiag_sym")
        # funm.py:136-145: any pytree `vec`, any callable `matvec(vec, *parameters)`
        op, v, unravel = _generic.wrap(matvec, vec, parameters)
        lib = _lib.load()
        n = v.shape[0]
        k = _check(op, n)
        known = _known()
        if _use_two_pass(op, known):
            return unravel(_two_pass_blocked(op, v.reshape(n, 1).contiguous(), 1, known)[:, 0])
        alphas, betas, init_len, Q, _ = decomp.lanczos_blocked(
            op, v.reshape(n, 1), k, spec["reortho"], want_Q=True, want_residual=False)
        ld = 1
        if known is not None:
            coeffs = torch.empty((k, ld), dtype=op.dtype, device=v.device)
            ws = _device.workspace(lib.mf_tridiag_quad_workspace_bytes(ld, k))
            _lib.check(lib.mf_tridiag_funm_e1(alphas.data_ptr(), betas.data_ptr(),
                                              _device.mf_dtype(op.dtype), ld, 1, k, known[0],
                                              known[1], coeffs.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _device.stream()))
        else:
            T = decomp._todense_tridiag_sym(alphas[:, 0], betas[: k - 1, 0])
            coeffs = _device.as_device(dense_funm(T), op.dtype)[:, 0].reshape(k, 1).contiguous()
        out = torch.empty((n, ld), dtype=op.dtype, device=v.device)
        Qc = Q if Q.is_contiguous() else Q.contiguous()
        _lib.check(lib.mf_basis_combine(Qc.data_ptr(), coeffs.data_ptr(), init_len.data_ptr(),
                                        _device.mf_dtype(op.dtype), n, ld, k, out.data_ptr(),
                                        _device.stream()))
        return unravel(out[:, 0])

    def batched(matvec, V, *, tile=None):
        """``f(A) V[p]`` for every row of ``V (P, n)``; returns ``(P, n)``."""
        import torch

        if spec is None:
            raise TypeError("funm_lanczos_sym: tridiag_sym must come from matfree_b200.decomp.tridiag_sym")
        op = ops.require_operator(matvec, "funm_lanczos_sym")
        lib = _lib.load()
        V = _device.as_device(V, op.dtype)
        P, n = V.shape
        _check(op, n)
        known = _known()
        if not _use_two_pass(op, known):
            return torch.stack([estimate(op, V[p]) for p in range(P)])
        ld = int(tile) if tile else _device.ld_for(P)
        mfdt = _device.mf_dtype(op.dtype)
        out = torch.empty_like(V)
        Xb = torch.zeros((n, ld), dtype=op.dtype, device=V.device)
        for p0 in range(0, P, ld):
            npb = min(ld, P - p0)
            if npb < ld:
                Xb.zero_()
            _lib.check(lib.mf_to_blocked(V[p0:p0 + npb].data_ptr(), Xb.data_ptr(), mfdt, n, npb, ld,
                                         _device.stream()))
            Wb = _two_pass_blocked(op, Xb, npb, known)
            _lib.check(lib.mf_from_blocked(Wb.data_ptr(), out[p0:p0 + npb].data_ptr(), mfdt, n, npb, ld,
                                           _device.stream()))
        return out

    def blocked(matvec, V0b, num_probes=None):
        """Blocked entry for callers that already hold the probes as ``V0b[n][ld]``: returns
        ``f(A) V0b`` in the same layout (two-pass route only)."""
        op = ops.require_operator(matvec, "funm_lanczos_sym")
        known = _known()
        if spec is None or not _use_two_pass(op, known):
            raise TypeError("funm_lanczos_sym.blocked needs reortho='none' and a recognised matfun")
        _check(op, V0b.shape[0])
        return _two_pass_blocked(op, V0b, V0b.shape[1] if num_probes is None else num_probes, known)

    estimate.batched = batched
    estimate.blocked = blocked
    return estimate


def funm_arnoldi(dense_funm, hessenberg, /):
    """Matrix-function-vector product ``f(A) v`` via the Arnoldi iteration, for arbitrary square
    operators (`funm.py:150-183`): ``|v| Q^T f(H) e1`` with `decomp.hessenberg`'s ``(Q, H)``.
    The small ``k x k`` function is `dense_funm` (`dense_funm_schur`, `dense_funm_pade_exp`, ...);
    the combination with the stored basis is `mf_basis_combine`."""

    def estimate(matvec, vec, *parameters):
        import torch

        from matfree_b200 import _generic
        from matfree_b200.backend import tree

        vec_flat, unravel = tree.ravel_pytree(vec, matvec.dtype if isinstance(matvec, ops.Operator) else None)
        length = torch.linalg.vector_norm(vec_flat)

        def matvec_flat(v_f, *p):
            return tree.ravel_pytree(matvec(unravel(v_f), *p))[0]

        mv = matvec if isinstance(matvec, ops.Operator) else matvec_flat
        basis, matrix, *_ = hessenberg(mv, vec_flat / length, *parameters)   # funm.py:177
        dt = vec_flat.dtype
        fH = _device.as_device(dense_funm(matrix), dt)                       # :178
        k, n = basis.shape
        coeffs = fH[:, 0].reshape(k, 1).contiguous()                          # f(H) e1, :179
        out = torch.empty((n, 1), dtype=dt, device=vec_flat.device)
        Q = basis.reshape(k, n, 1).contiguous()
        scale = length.reshape(1).to(dt).contiguous()
        _lib.check(_lib.load().mf_basis_combine(Q.data_ptr(), coeffs.data_ptr(), scale.data_ptr(),
                                                _device.mf_dtype(dt), n, 1, k, out.data_ptr(),
                                                _device.stream()))
        return unravel(out[:, 0])                                             # :181

    return estimate


def dense_funm_schur(matfun):
    """Dense matrix function of a (possibly non-symmetric) small matrix (`funm.py:338-347`,
    `linalg.funm_schur` = `jax.scipy.linalg.funm`: Schur-Parlett).  The ``k x k`` factor is
    evaluated on the host with SciPy's Schur-Parlett `funm`, like the reference's LAPACK call."""

    def fun(dense_matrix):
        import scipy.linalg
        import torch

        M = _device.as_device(dense_matrix)
        H = M.detach().cpu().numpy().astype(np.float64)

        def f(x):
            return np.asarray(_apply_matfun_host(matfun, x))

        out = scipy.linalg.funm(H, f, disp=True)
        return torch.as_tensor(np.real(out), device=M.device).to(M.dtype)

    fun._mf_matfun = matfun
    return fun


def _apply_matfun_host(matfun, x):
    known = _known_fn(matfun) if not callable(matfun) or _is_hashable(matfun) else None
    if known is not None:
        fn, param = known
        return {_lib.MF_FN_LOG: np.log, _lib.MF_FN_EXP: lambda t: np.exp(param * t),
                _lib.MF_FN_INV: lambda t: 1.0 / t, _lib.MF_FN_SQRT: np.sqrt,
                _lib.MF_FN_POW: lambda t: t ** param, _lib.MF_FN_IDENTITY: lambda t: t,
                _lib.MF_FN_SIN: lambda t: np.sin(param * t)}[fn](x)
    try:
        return np.asarray(matfun(x))
    except Exception:
        import torch

        return matfun(torch.as_tensor(x)).numpy()


def dense_funm_pade_exp():
    """Dense matrix exponential by a Pade approximation (`funm.py:350-359`,
    `jax.scipy.linalg.expm`); `torch.linalg.matrix_exp` on the small ``k x k`` factor."""

    def fun(dense_matrix):
        import torch

        return torch.linalg.matrix_exp(_device.as_device(dense_matrix))

    return fun


# ---------------------------------------------------------------- products A^T A (bidiag)


def dense_funm_product_svd(matfun):
    """Dense matrix function of ``B^T B`` via the SVD of ``B`` (`funm.py:305-319`)."""

    def dense_funm(matrix, /):
        import torch

        M = _device.as_device(matrix)
        _, S, Vt = torch.linalg.svd(M, full_matrices=False)
        fx = _apply_matfun(matfun, S**2)
        return Vt.T @ (fx[:, None] * Vt)

    dense_funm._mf_matfun = matfun
    return dense_funm


def product_quadrature_blocked(alphas, betas, init_len, num_probes, matfun):
    """``init_len^2 * e1^T f(B^T B) e1`` for every probe of a tile (`mf_bidiag_quad`)."""
    import torch

    lib = _lib.load()
    k, ld = alphas.shape
    dt, dev = alphas.dtype, alphas.device
    ws = _device.workspace(lib.mf_tridiag_quad_workspace_bytes(ld, k))
    known = _known_fn(matfun) if (not callable(matfun) or _is_hashable(matfun)) else None
    if known is not None:
        quad = torch.empty((ld,), dtype=dt, device=dev)
        _lib.check(lib.mf_bidiag_quad(alphas.data_ptr(), betas.data_ptr(), init_len.data_ptr(),
                                      _device.mf_dtype(dt), ld, num_probes, k, known[0], known[1],
                                      quad.data_ptr(), None, None, ws.data_ptr(), ws.numel(),
                                      _device.stream()))
        return quad[:num_probes]
    nodes = torch.empty((k, ld), dtype=torch.float64, device=dev)
    weights = torch.empty((k, ld), dtype=torch.float64, device=dev)
    _lib.check(lib.mf_bidiag_quad(alphas.data_ptr(), betas.data_ptr(), init_len.data_ptr(),
                                  _device.mf_dtype(dt), ld, num_probes, k, _lib.MF_FN_NONE, 0.0, None,
                                  nodes.data_ptr(), weights.data_ptr(), ws.data_ptr(), ws.numel(),
                                  _device.stream()))
    fx = _apply_matfun(matfun, nodes[:, :num_probes].to(dt)).to(torch.float64)
    q = (fx * weights[:, :num_probes]).sum(dim=0) * init_len[:num_probes].to(torch.float64) ** 2
    return q.to(dt)


def monte_carlo_funm_product(dense_funm, bidiag, /):
    """Integrand for the trace of a function of ``A^T A`` (`funm.py:275-302`)."""
    spec = getattr(bidiag, "_mf_spec", None)
    matfun = getattr(dense_funm, "_mf_matfun", None)
    fusable = spec is not None and spec.get("kind") == "bidiag" and matfun is not None

    def quadform(matvec, v0, *parameters):
        import torch

        from matfree_b200.backend import tree

        if not fusable or not isinstance(matvec, ops.Operator):
            # funm.py:287-300 with user-supplied pieces: pytree in, (possibly different) pytree out
            v, unravel = tree.ravel_pytree(v0)
            length = torch.linalg.vector_norm(v)

            def matvec_flat(v_f, *p):
                return tree.ravel_pytree(matvec(unravel(v_f), *p))[0]

            mv = matvec if isinstance(matvec, ops.Operator) else matvec_flat
            _, B, *_ = bidiag(mv, v / length, *parameters)
            return length**2 * dense_funm(B)[0, 0]
        if parameters:
            raise TypeError("registered operators carry their own buffers; extra matvec parameters "
                            "are only supported for callables")
        op = matvec
        if not isinstance(op, ops.RectOperator):
            raise TypeError("monte_carlo_funm_product: a registered matvec must be ops.rect(A)")
        v = _device.as_device(v0, op.dtype).reshape(-1)
        k = spec["num_matvecs"]
        if k > min(op.m, op.n) or k < 0:
            raise ValueError(decomp._error_num_matvecs(k, maxval=min(op.m, op.n), minval=0))
        alphas, betas, init_len, *_ = decomp.bidiag_blocked(op, v.reshape(-1, 1).contiguous(), k, spec["reortho"])
        return product_quadrature_blocked(alphas, betas, init_len, 1, matfun)[0]

    quadform._mf_integrand = None if not fusable else {
        "kind": "product", "num_matvecs": spec["num_matvecs"], "reortho": spec["reortho"], "matfun": matfun}
    return quadform


def monte_carlo_funm_product_logdet(bidiag, /):
    """Integrand for ``logdet(A^T A)`` (`funm.py:246-255`)."""
    return monte_carlo_funm_product(dense_funm_product_svd(np.log), bidiag)


def monte_carlo_funm_product_schatten_norm(power, bidiag, /):
    """Integrand for the p-th power of the Schatten-p norm (`funm.py:258-272`):
    ``f(x) = x^(p/2)`` applied to the eigenvalues of ``A^T A``."""
    return monte_carlo_funm_product(dense_funm_product_svd(("pow", power / 2)), bidiag)

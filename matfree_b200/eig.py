"""Partial eigen- and singular-value decompositions -- mirrors `matfree/eig.py:22-160`
(`svd_partial`, `eigh_partial`, `eig_partial`): thin consumers of `decomp.tridiag_sym` /
`decomp.hessenberg` / `decomp.bidiag`; vectors may be pytrees, matvecs callables with parameters.

The Krylov bases stay on the device; the small ``k x k`` factor is decomposed with LAPACK through
torch (k is the number of matvecs: tens), and the Ritz / singular vectors ``S^T Q`` are formed by
the library's basis-combination kernel (`mf_basis_combine`: one pass over the stored basis per
vector).
"""

from __future__ import annotations

from matfree_b200 import _device, _lib


def _combine_rows(S, Q):
    """``S @ Q`` for a small ``S (r, k)`` and a basis ``Q (k, n)`` whose rows are Krylov vectors."""
    import torch

    lib = _lib.load()
    r, k = S.shape
    n = Q.shape[1]
    out = torch.empty((r, n), dtype=Q.dtype, device=Q.device)
    if r == 0 or k == 0:
        return out
    Qb = Q.contiguous().reshape(k, n, 1)  # blocked basis [k][n][ld = 1]
    S = S.to(Q.dtype).contiguous()
    for i in range(r):
        _lib.check(lib.mf_basis_combine(Qb.data_ptr(), S[i].reshape(k, 1).contiguous().data_ptr(), None,
                                        _device.mf_dtype(Q.dtype), n, 1, k, out[i].data_ptr(),
                                        _device.stream()))
    return out


def _partial_and_flatten_matvec(Av, v0, parameters):
    """`eig.py:145-160`: bind the parameters and conjugate the matvec with ravel / unravel.  A
    registered operator already takes flat vectors (and carries its buffers)."""
    from matfree_b200 import ops
    from matfree_b200.backend import tree

    if isinstance(Av, ops.Operator):
        if parameters:
            raise TypeError("registered operators carry their own buffers; extra matvec parameters "
                            "are only supported for callables")
        v0_flat, v_unravel = tree.ravel_pytree(v0, Av.dtype)
        return Av, v0_flat, v_unravel, None
    v0_flat, v_unravel = tree.ravel_pytree(v0)
    holder = {}

    def Av_flat(v_flat):
        result_flat, holder["u_unravel"] = tree.ravel_pytree(Av(v_unravel(v_flat), *parameters), v_flat.dtype)
        return result_flat

    return Av_flat, v0_flat, v_unravel, holder


def eigh_partial(tridiag_sym):
    """Partial eigendecomposition ``A ~ V diag(vals) V^T`` of a symmetric operator
    (`eig.py:69-104`): returns ``(vals (k,), vecs (k, n))``, rows of `vecs` are Ritz vectors."""

    def eigh(Av, v0, *parameters):
        import torch

        Av_flat, v0_flat, v_unravel, _ = _partial_and_flatten_matvec(Av, v0, parameters)
        Q, H, *_ = tridiag_sym(Av_flat, v0_flat)
        if isinstance(H, tuple):  # materialize=False: (diag, offdiag)
            from matfree_b200 import decomp

            H = decomp._todense_tridiag_sym(*H)
        vals, vecs = torch.linalg.eigh(H)
        return vals, v_unravel.batched(_combine_rows(vecs.T, Q))

    return eigh


def eig_partial(hessenberg):
    """Partial eigendecomposition of an arbitrary square operator via the Hessenberg
    factorisation (`eig.py:107-142`): ``vals, vecs = eig(H); vecs = vecs^T Q`` -- complex in
    general (`jax.numpy.linalg.eig`).  The ``k x k`` eigenproblem is LAPACK's (`torch.linalg.eig`),
    the real and imaginary parts of ``vecs^T Q`` are two passes of `mf_basis_combine` over the
    stored basis."""

    def eig(Av, v0, *parameters):
        import torch

        Av_flat, v0_flat, v_unravel, _ = _partial_and_flatten_matvec(Av, v0, parameters)
        Q, H, *_ = hessenberg(Av_flat, v0_flat)
        vals, vecs = torch.linalg.eig(H)
        St = vecs.T.contiguous()
        out = torch.complex(_combine_rows(St.real.contiguous(), Q), _combine_rows(St.imag.contiguous(), Q))
        return vals, v_unravel.batched(out)

    return eig


def svd_partial(bidiag):
    """Partial SVD ``A ~ U^T diag(s) V`` via bidiagonalisation (`eig.py:22-66`): returns
    ``(ut (k, nrows), s (k,), vt (k, ncols))``.  Assumes the bidiagonalisation materialises B."""

    def svd(Av, v0, *parameters):
        import torch

        Av_flat, v0_flat, v_unravel, holder = _partial_and_flatten_matvec(Av, v0, parameters)
        (u, v), B, *_ = bidiag(Av_flat, v0_flat)
        if isinstance(B, tuple):
            raise TypeError("svd_partial assumes that the bidiagonalisation materialises the bidiagonal matrix")
        U, S, Vt = torch.linalg.svd(B, full_matrices=False)
        ut, vt = _combine_rows(U.T, u), _combine_rows(Vt, v)
        if holder and "u_unravel" in holder:
            ut = holder["u_unravel"].batched(ut)
        return ut, S, v_unravel.batched(vt)

    return svd

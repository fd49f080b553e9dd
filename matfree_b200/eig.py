"""Partial eigen- and singular-value decompositions -- mirrors `matfree/eig.py:22-104`
(`svd_partial`, `eigh_partial`): thin consumers of `decomp.tridiag_sym` / `decomp.bidiag`.

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


def eigh_partial(tridiag_sym):
    """Partial eigendecomposition ``A ~ V diag(vals) V^T`` of a symmetric operator
    (`eig.py:69-104`): returns ``(vals (k,), vecs (k, n))``, rows of `vecs` are Ritz vectors."""

    def eigh(Av, v0, *parameters):
        import torch

        Q, H, *_ = tridiag_sym(Av, v0, *parameters)
        if isinstance(H, tuple):  # materialize=False: (diag, offdiag)
            from matfree_b200 import decomp

            H = decomp._todense_tridiag_sym(*H)
        vals, vecs = torch.linalg.eigh(H)
        return vals, _combine_rows(vecs.T, Q)

    return eigh


def svd_partial(bidiag):
    """Partial SVD ``A ~ U^T diag(s) V`` via bidiagonalisation (`eig.py:22-66`): returns
    ``(ut (k, nrows), s (k,), vt (k, ncols))``.  Assumes the bidiagonalisation materialises B."""

    def svd(Av, v0, *parameters):
        import torch

        (u, v), B, *_ = bidiag(Av, v0, *parameters)
        if isinstance(B, tuple):
            raise TypeError("svd_partial assumes that the bidiagonalisation materialises the bidiagonal matrix")
        U, S, Vt = torch.linalg.svd(B, full_matrices=False)
        return _combine_rows(U.T, u), S, _combine_rows(Vt, v)

    return svd

"""Registered matvec operators: dense, CSR, Gram.

The reference takes any JAX-traceable callable ``matvec(v, *params) -> Av``
(`matfree/stochtrace.py:47-49`, `matfree/funm.py:231-235`,
`matfree/decomp.py:163-164`).  The operators below keep that callable
signature -- ``op(v)`` returns ``A @ v`` for a flat device vector -- and
additionally carry their device buffers, so `decomp`, `funm` and `stochtrace`
can hand the whole probe block to the fused CUDA kernels
(`mf_matmat_*`, `mf_lanczos`, `mf_estimate` in `include/matfree_b200.h`).
A callable that is not one of these raises in the fused entry points: there is
no CPU (or generic-callable) fallback for the Lanczos kernels.
"""

from __future__ import annotations

import ctypes

import numpy as np

from matfree_b200 import _device, _lib


class Operator:
    """Base class of the registered operators (callable: ``op(v, *params)``)."""

    kind: int
    n: int
    dtype = None  # torch dtype

    def _struct(self) -> _lib.MfOperator:
        raise NotImplementedError

    @property
    def shape(self):
        return (self.n, self.n)

    def _make_planes(self):
        """TF32 hi/lo planes of an fp32 dense / Gram matrix (`mf_operator_split`): with
        them the operator runs on the tcgen05 tensor cores (3xTF32)."""
        lib = _lib.load()
        self._planes = None
        st = self._struct()
        nbytes = lib.mf_operator_split_bytes(ctypes.byref(st))
        if nbytes <= 0:
            return
        planes = _device.workspace(nbytes)
        _lib.check(lib.mf_operator_split(ctypes.byref(st), planes.data_ptr(), _device.stream()))
        self._planes = planes

    # -- the reference's callable signature
    def __call__(self, v, *params):
        if params:
            raise TypeError("registered operators carry their own buffers; extra matvec parameters are not supported")
        v = _device.as_device(v, self.dtype)
        if v.ndim != 1 or v.shape[0] != self.n:
            raise ValueError(f"expected a flat vector of length {self.n}, got shape {tuple(v.shape)}")
        return self.matmat_blocked(v.reshape(self.n, 1)).reshape(self.n)

    def matmat_blocked(self, X):
        """``A @ X`` for a blocked device array ``X[n][ld]`` (ld a power of two <= 256)."""
        import torch

        lib = _lib.load()
        n, ld = X.shape
        assert n == self.n and X.is_contiguous() and X.dtype == self.dtype
        W = torch.empty_like(X)
        st = self._struct()
        nbytes = lib.mf_matmat_workspace_bytes(ctypes.byref(st), ld)
        if nbytes < 0:
            _lib.check(-1)
        ws = _device.workspace(nbytes)
        _lib.check(lib.mf_matmat(ctypes.byref(st), X.data_ptr(), W.data_ptr(), ld, ws.data_ptr(),
                                 ws.numel(), _device.stream()))
        return W

    def matmat(self, V):
        """``V @ A^T`` for probe-major ``V (P, n)``; returns ``(P, n)`` (any P)."""
        import torch

        lib = _lib.load()
        V = _device.as_device(V, self.dtype)
        P, n = V.shape
        out = torch.empty_like(V)
        ld = _device.ld_for(P)
        Xb = torch.empty((n, ld), dtype=self.dtype, device=V.device)
        mfdt = _device.mf_dtype(self.dtype)
        for p0 in range(0, P, ld):
            npb = min(ld, P - p0)
            _lib.check(lib.mf_to_blocked(V[p0:p0 + npb].data_ptr(), Xb.data_ptr(), mfdt, n, npb, ld,
                                         _device.stream()))
            Wb = self.matmat_blocked(Xb)
            _lib.check(lib.mf_from_blocked(Wb.data_ptr(), out[p0:p0 + npb].data_ptr(), mfdt, n, npb,
                                           ld, _device.stream()))
        return out


class DenseOperator(Operator):
    kind = _lib.MF_OP_DENSE

    def __init__(self, A):
        A = _device.as_device(A)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("ops.dense expects a square matrix")
        self.dtype = _device.torch_dtype(A.dtype)
        self.A = A
        self.n = int(A.shape[0])
        self._planes = None
        self._make_planes()

    def _struct(self):
        return _lib.MfOperator(kind=self.kind, dtype=_device.mf_dtype(self.dtype), n=self.n, m=self.n,
                               nnz=0, values=self.A.data_ptr(), indptr=None, indices=None,
                               lda=self.n,
                               split_planes=None if self._planes is None else self._planes.data_ptr())


class CsrOperator(Operator):
    kind = _lib.MF_OP_CSR

    def __init__(self, indptr, indices, data, n=None):
        import torch

        self.data = _device.as_device(data)
        self.dtype = _device.torch_dtype(self.data.dtype)
        self.indptr = _device.as_device(indptr, torch.int32)
        self.indices = _device.as_device(indices, torch.int32)
        self.n = int(self.indptr.shape[0] - 1) if n is None else int(n)
        if self.indptr.shape[0] != self.n + 1:
            raise ValueError("ops.csr: indptr must have n + 1 entries")
        self.nnz = int(self.data.shape[0])
        if self.indices.shape[0] != self.nnz:
            raise ValueError("ops.csr: indices and data must have the same length")
        # longest row: above 128 non-zeros the SpMM takes its load-balanced route
        counts = self.indptr[1:] - self.indptr[:-1]
        self.max_row_nnz = int(counts.max()) if self.n > 0 else 0
        # bandwidth max |col - row| from the first / last entry of every row (columns ascending):
        # lets the product pick a blocked row order when planes of a 3-D stencil exceed the L2
        self.bandwidth = 0
        if self.nnz > 0:
            rows = torch.arange(self.n, device=self.indptr.device, dtype=torch.int64)
            ne = counts > 0
            first = self.indices[self.indptr[:-1][ne].long()].long()
            last = self.indices[(self.indptr[1:][ne] - 1).long()].long()
            self.bandwidth = int(torch.maximum((first - rows[ne]).abs(), (last - rows[ne]).abs()).max())
        self.num_diagonals = self._count_diagonals(counts)

    def _count_diagonals(self, counts):
        """Number of distinct diagonals (col - row) if every entry lies on the diagonals of the longest
        row, that row has at most 8 entries and its offsets look like a grid stencil's (symmetric, with
        -1 / 0 / +1), else 255 ("many / not a stencil"); 0 for an empty matrix.  A hint for the
        kernel choice (`mf_operator_t::csr_num_diagonals`): stencil matrices qualify, irregular ones do not."""
        import torch

        self.line_stride = 0
        if self.nnz == 0 or self.n == 0:
            return 0
        if self.max_row_nnz > 8 or self.nnz >= 2 ** 31:
            return 255
        k = int(torch.argmax(counts))
        j0 = int(self.indptr[k])
        offs = (self.indices[j0:j0 + self.max_row_nnz] - k).tolist()
        # a grid stencil: distinct, symmetric offsets around the three adjacent middle diagonals
        if len(set(offs)) != len(offs) or sorted(offs) != sorted(-o for o in offs) or not {-1, 0, 1} <= set(offs):
            return 255
        row_of = torch.repeat_interleave(torch.arange(self.n, device=self.indices.device, dtype=torch.int32),
                                         counts.to(torch.int64))
        d = self.indices - row_of
        del row_of
        on = torch.zeros_like(d, dtype=torch.bool)
        for o in offs:
            on |= d == o
        if not bool(on.all()):
            return 255
        above = [o for o in offs if o > 1]
        self.line_stride = min(above) if above else 0  # `mf_operator_t::csr_line_stride`
        return len(offs)

    def _struct(self):
        return _lib.MfOperator(kind=self.kind, dtype=_device.mf_dtype(self.dtype), n=self.n, m=self.n,
                               nnz=self.nnz, values=self.data.data_ptr(),
                               indptr=self.indptr.data_ptr(), indices=self.indices.data_ptr(),
                               lda=0, split_planes=None, csr_max_row_nnz=self.max_row_nnz,
                               csr_bandwidth=self.bandwidth, csr_num_diagonals=self.num_diagonals,
                               csr_line_stride=self.line_stride)


class GramOperator(Operator):
    """``v -> A^T (A v)`` for a rectangular ``A (m, n)`` (tutorial 1's operator)."""

    kind = _lib.MF_OP_GRAM

    def __init__(self, A):
        A = _device.as_device(A)
        if A.ndim != 2:
            raise ValueError("ops.gram expects a matrix")
        self.dtype = _device.torch_dtype(A.dtype)
        self.A = A
        self.m, self.n = int(A.shape[0]), int(A.shape[1])
        self._planes = None
        self._make_planes()

    def _struct(self):
        return _lib.MfOperator(kind=self.kind, dtype=_device.mf_dtype(self.dtype), n=self.n, m=self.m,
                               nnz=0, values=self.A.data_ptr(), indptr=None, indices=None,
                               lda=self.n,
                               split_planes=None if self._planes is None else self._planes.data_ptr())


class RectOperator(GramOperator):
    """``v -> A v`` for a rectangular dense ``A (m, n)`` -- the (non-symmetric) matvec the
    Golub-Kahan bidiagonalisation takes (`matfree/decomp.py:608-750`); the reference obtains
    ``u -> A^T u`` from it with `jax.vjp` (`decomp.py:703,712`), here both products are one GEMM
    each on the operator's buffers (`mf_matmat_rect`).  Shares the storage (and, for fp32, the
    TF32 planes) of the Gram operator over the same matrix."""

    def __call__(self, v, *params):
        if params:
            raise TypeError("registered operators carry their own buffers; extra matvec parameters are not supported")
        v = _device.as_device(v, self.dtype)
        if v.ndim != 1 or v.shape[0] != self.n:
            raise ValueError(f"expected a flat vector of length {self.n}, got shape {tuple(v.shape)}")
        return self.apply_blocked(v.reshape(self.n, 1), trans=False).reshape(self.m)

    @property
    def shape(self):
        return (self.m, self.n)

    def apply_blocked(self, X, *, trans: bool):
        """``A @ X`` (``X[n][ld] -> [m][ld]``) or ``A^T @ X`` (``X[m][ld] -> [n][ld]``)."""
        import torch

        lib = _lib.load()
        rows_in, rows_out = (self.m, self.n) if trans else (self.n, self.m)
        k, ld = X.shape
        assert k == rows_in and X.is_contiguous() and X.dtype == self.dtype
        W = torch.empty((rows_out, ld), dtype=self.dtype, device=X.device)
        mfdt = _device.mf_dtype(self.dtype)
        have = self._planes is not None
        ws = _device.workspace(lib.mf_matmat_rect_workspace_bytes(self.m, self.n, self.n, int(trans), mfdt,
                                                                  int(have), ld))
        _lib.check(lib.mf_matmat_rect(self.A.data_ptr(), self._planes.data_ptr() if have else None,
                                      self.m, self.n, self.n, int(trans), mfdt, X.data_ptr(),
                                      W.data_ptr(), ld, ws.data_ptr(), ws.numel(), _device.stream()))
        return W

    def matmat_blocked(self, X):
        return self.apply_blocked(X, trans=False)

    def rmatmat_blocked(self, U):
        return self.apply_blocked(U, trans=True)

    def matmat(self, V):
        raise TypeError("ops.rect is not square: use apply_blocked")


def rect(A) -> RectOperator:
    """Rectangular dense operator ``v -> A @ v`` (for `decomp.bidiag` and the
    `funm.monte_carlo_funm_product*` integrands)."""
    return RectOperator(A)


def dense(A) -> DenseOperator:
    """Symmetric dense operator ``v -> A @ v``."""
    return DenseOperator(A)


def csr(indptr, indices, data, n=None) -> CsrOperator:
    """CSR operator (int32 ``indptr``/``indices``)."""
    return CsrOperator(indptr, indices, data, n)


def csr_from_scipy(mat, dtype=None) -> CsrOperator:
    mat = mat.tocsr()
    data = mat.data if dtype is None else mat.data.astype(dtype)
    return CsrOperator(mat.indptr.astype(np.int32), mat.indices.astype(np.int32), data, mat.shape[0])


def csr_row_sharded(indptr, indices, data, n, row_start, group=None):
    """Rows ``[row_start, row_start + len(indptr) - 1)`` of a global ``n x n`` CSR operator; the
    other rows live on the other ranks of `group` (`matfree_b200._rowshard`).  `indptr` is local
    (starts at 0), `indices` are global column ids.  Vectors passed to / returned by the
    decompositions are this rank's row slab."""
    from matfree_b200 import _rowshard

    return _rowshard.RowShardedCsr(indptr, indices, data, n, row_start, group)


def gram(A) -> GramOperator:
    """Gram operator ``v -> A^T (A v)``."""
    return GramOperator(A)


def require_operator(matvec, who: str) -> Operator:
    if not isinstance(matvec, Operator):
        raise TypeError(
            f"{who}: matvec must be a registered operator (matfree_b200.ops.dense / csr / gram); "
            f"got {type(matvec).__name__}. The Lanczos kernels run on the GPU and have no "
            "generic-callable or CPU fallback."
        )
    return matvec

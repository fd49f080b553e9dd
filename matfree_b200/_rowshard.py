"""Row-sharded operators and decompositions (one process per GPU, `torch.distributed`).

For operators too large for one GPU, and for BASELINE config 4 (`tridiag_sym` with full
re-orthogonalisation on a 256^3 Laplacian over 8 GPUs), the rows of the operator -- and with
them every Lanczos vector -- are partitioned into contiguous slabs, one per rank:

* matvec: each rank owns the CSR rows ``[r0, r1)``; the columns it touches outside its slab
  (the *halo*: one grid plane per side for a stencil) are fetched from the owning ranks before
  every product (`HaloPlan`, NCCL send/recv over NVLink; for a matrix without band structure
  the plan degenerates to an all-gather of the block);
* reductions: the dot products / norms of `matfree/decomp.py:454-477` (Arnoldi + CGS twice) and
  `:286-292` (three-term step) become per-rank fp64 partial sums (`mf_reorth_dots`,
  `mf_block_dot`, ... in `include/matfree_b200.h`) followed by one all-reduce each.

The step loop lives here (collectives between kernels); all arithmetic on vectors is in the
CUDA library.  `Backend` is the seam the CPU `gloo` tests use to exercise this file's logic
(partitioning, halo plan, collective placement) without a GPU.
"""

from __future__ import annotations

import ctypes

import numpy as np

from matfree_b200 import _device, _lib, ops


# ----------------------------------------------------------------------------- partitioning


def slab_range(n: int, world: int, rank: int, align: int = 1):
    """Contiguous row slab ``[r0, r1)`` of rank `rank`; slab boundaries are multiples of
    `align` (e.g. one grid plane) so that a stencil's halo is whole planes."""
    units = -(-n // align)
    per = -(-units // world)
    r0 = min(n, rank * per * align)
    r1 = min(n, (rank + 1) * per * align)
    return r0, r1


class HaloPlan:
    """Who sends which rows to whom so that every rank holds the columns its rows touch.

    `ranges[r] = (r0, r1)` are the owned slabs, `needs[r] = (c0, c1)` the half-open column range
    rank r's rows reference (c0 <= r0, c1 >= r1).  The extended local block of rank r covers rows
    ``[c0, c1)``; `sends` / `recvs` list ``(peer, a, b)`` global row ranges.
    """

    def __init__(self, rank: int, ranges, needs):
        self.rank = rank
        self.ranges = [tuple(int(x) for x in r) for r in ranges]
        self.needs = [tuple(int(x) for x in c) for c in needs]
        r0, r1 = self.ranges[rank]
        c0, c1 = self.needs[rank]
        self.r0, self.r1, self.c0, self.c1 = r0, r1, min(c0, r0), max(c1, r1)
        self.lo = self.r0 - self.c0           # halo rows below the slab
        self.hi = self.c1 - self.r1           # halo rows above the slab
        self.n_ext = self.c1 - self.c0
        # unused leading rows of the extended block so that the slab starts on a 4-row
        # boundary: keeps 16-byte vector accesses legal for narrow tiles (ld = 1, 2)
        self.pad = (-self.lo) % 4
        self.recvs, self.sends = [], []
        for peer, (p0, p1) in enumerate(self.ranges):
            if peer == rank:
                continue
            # what I need from peer: my column range intersected with its slab, outside my slab
            a, b = max(self.c0, p0), min(self.c1, p1)
            if a < b:
                self.recvs.append((peer, a, b))
            # what peer needs from me
            q0, q1 = self.needs[peer]
            a, b = max(min(q0, p0), r0), min(max(q1, p1), r1)
            if a < b:
                self.sends.append((peer, a, b))

    @property
    def halo_rows(self):
        return sum(b - a for _, a, b in self.recvs)

    @property
    def rows_alloc(self):
        """Rows of the extended block as allocated (padding + halo + slab + halo)."""
        return (self.pad + self.n_ext + 3) // 4 * 4

    def row(self, a: int) -> int:
        """Index of global row `a` in the extended block."""
        return self.pad + a - self.c0


    def c_struct(self):
        """`mf_halo_plan_t` of this rank (peer-memory exchange): every send names the row of the
        RECEIVER's extended block it lands in, which this rank can compute because all slabs and
        column ranges are known to everybody."""
        if getattr(self, "_c", None) is None:
            peers = {}
            sends = (_lib.MfHaloSend * max(len(self.sends), 1))()
            for i, (peer, a, b) in enumerate(self.sends):
                pp = peers.get(peer) or peers.setdefault(peer, HaloPlan(peer, self.ranges, self.needs))
                sends[i] = _lib.MfHaloSend(peer=peer, src_row=self.row(a), rows=b - a,
                                           dst_row=pp.row(a), dst_rows_alloc=pp.rows_alloc)
            srcs = sorted({peer for peer, _, _ in self.recvs})
            recv = (ctypes.c_int32 * max(len(srcs), 1))(*srcs)
            plan = _lib.MfHaloPlan(rows_alloc=self.rows_alloc, mid_row=self.row(self.r0),
                                   num_sends=len(self.sends), sends=sends,
                                   num_recv_peers=len(srcs), recv_peers=recv)
            self._c = (plan, sends, recv)  # keep the arrays alive
        return self._c[0]


def make_plan(r0: int, r1: int, cmin: int, cmax_excl: int, group=None) -> HaloPlan:
    """Exchange ``(r0, r1, cmin, cmax)`` between the ranks of `group` and build the plan."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return HaloPlan(0, [(r0, r1)], [(cmin, cmax_excl)])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = _device.device() if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([r0, r1, cmin, cmax_excl], dtype=torch.int64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    rows = [(int(t[0]), int(t[1])) for t in allv]
    cols = [(int(t[2]), int(t[3])) for t in allv]
    return HaloPlan(rank, rows, cols)


def exchange_halo(plan: HaloPlan, Xext, group=None):
    """Fill the halo rows of the extended block `Xext[n_ext][ld]` (its middle already holds this
    rank's rows) from the owning ranks.  Rows of a blocked vector are contiguous, so every
    message is one contiguous slice."""
    import torch.distributed as dist

    if not plan.sends and not plan.recvs:
        return
    ops_ = []
    for peer, a, b in plan.recvs:
        ops_.append(dist.P2POp(dist.irecv, Xext[plan.row(a):plan.row(b)], _global_rank(peer, group), group))
    for peer, a, b in plan.sends:
        ops_.append(dist.P2POp(dist.isend, Xext[plan.row(a):plan.row(b)], _global_rank(peer, group), group))
    for req in dist.batch_isend_irecv(ops_):
        req.wait()


def _global_rank(group_rank, group):
    import torch.distributed as dist

    if group is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


def _all_reduce(t, group):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)


# ----------------------------------------------------------------------------- peer memory


class _DevView:
    """Raw device memory as a `__cuda_array_interface__` object (for `torch.as_tensor`)."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}
        self._owner = owner


class PeerComm:
    """`mf_comm_t` for the ranks of a process group: one peer-mapped region per rank (CUDA IPC),
    over which the library's kernels exchange halos and all-reduce their sums (NVLink loads /
    stores, no NCCL on the per-step path).  torch.distributed only carries the 64-byte handles."""

    def __init__(self, group=None, heap_bytes=0):
        import torch
        import torch.distributed as dist

        self.lib = _lib.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.mf_comm_create(self.world, self.rank, int(heap_bytes), ctypes.byref(self._h)))
        buf = (ctypes.c_ubyte * _lib.MF_COMM_HANDLE_BYTES)()
        _lib.check(self.lib.mf_comm_handle(self._h, buf))
        dev = _device.device()
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        raw = bytes(torch.stack(allh).cpu().numpy().tobytes())
        _lib.check(self.lib.mf_comm_connect(self._h, raw))
        dist.barrier(group=group)  # every rank has mapped every region before anyone uses it
        self.heap_bytes = int(self.lib.mf_comm_heap_bytes(self._h))
        self._heap_ptr = int(self.lib.mf_comm_heap(self._h) or 0)

    @property
    def handle(self):
        return self._h

    def heap_view(self, offset, shape, dtype):
        """Tensor view of the local heap (the memory stays owned by the communicator)."""
        import torch

        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        assert offset + nbytes <= self.heap_bytes
        flat = torch.as_tensor(_DevView(self._heap_ptr + offset, nbytes, self), device=_device.device())
        return flat.view(dtype).view(*shape)

    def check(self):
        """Synchronise and raise if an in-kernel wait timed out."""
        _lib.check(self.lib.mf_comm_status(self._h, _device.stream()))

    def close(self, collective=True):
        """Unmap the peers' regions, wait until every rank has done so, free the own region.
        `collective=False` (interpreter shutdown) skips the rendezvous."""
        if self._h:
            import torch
            import torch.distributed as dist

            torch.cuda.synchronize()
            self.lib.mf_comm_disconnect(self._h)
            if collective and dist.is_available() and dist.is_initialized():
                dist.barrier(group=self.group)
            self.lib.mf_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass


def _use_peer_memory(group) -> bool:
    """Peer-memory drivers need one CUDA process per rank on one node (NCCL group)."""
    import os

    import torch.distributed as dist

    if os.environ.get("MF_ROWSHARD_NCCL"):
        return False  # cross-check route: Python step loop + NCCL collectives
    if not (dist.is_available() and dist.is_initialized()):
        return True  # single process: native driver without a communicator
    if dist.get_world_size(group) == 1:
        return True
    return dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= 8


# ----------------------------------------------------------------------------- operator


class RowShardedCsr(ops.Operator):
    """Rows ``[r0, r1)`` of a global ``n x n`` CSR operator (int32 `indptr` local, `indices`
    GLOBAL column ids).  The callable signature is kept: ``op(v_local)`` returns this rank's rows
    of ``A v`` (a collective call: every rank of the group must make it)."""

    kind = _lib.MF_OP_CSR

    def __init__(self, indptr, indices, data, n, row_start, group=None):
        import torch

        self.group = group
        self.data = _device.as_device(data)
        self.dtype = _device.torch_dtype(self.data.dtype)
        self.indptr = _device.as_device(indptr, torch.int32)
        gidx = _device.as_device(indices, torch.int64)
        self.n_global = int(n)
        self.n = int(self.indptr.shape[0] - 1)  # local rows
        self.r0 = int(row_start)
        self.r1 = self.r0 + self.n
        self.nnz = int(self.data.shape[0])
        if gidx.shape[0] != self.nnz:
            raise ValueError("ops.csr_row_sharded: indices and data must have the same length")
        if self.nnz:
            cmin, cmax = int(gidx.min()), int(gidx.max()) + 1
        else:
            cmin, cmax = self.r0, self.r1
        self.plan = make_plan(self.r0, self.r1, min(cmin, self.r0), max(cmax, self.r1), group)
        self.indices = (gidx + (self.plan.pad - self.plan.c0)).to(torch.int32)  # extended-block columns
        self.max_row_nnz = int((self.indptr[1:] - self.indptr[:-1]).max()) if self.n > 0 else 0

    def _struct(self):
        return _lib.MfOperator(kind=self.kind, dtype=_device.mf_dtype(self.dtype), n=self.n, m=self.n,
                               nnz=self.nnz, values=self.data.data_ptr(),
                               indptr=self.indptr.data_ptr(), indices=self.indices.data_ptr(),
                               lda=0, split_planes=None, csr_max_row_nnz=self.max_row_nnz)

    @property
    def shape(self):
        return (self.n_global, self.n_global)

    def peer_comm(self, need_bytes):
        """The operator's communicator with a heap of at least `need_bytes` (collective: the
        heap is (re)allocated when any rank needs more; sizes are cached per call signature)."""
        import torch
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return None
        comm = getattr(self, "_comm", None)
        key = int(need_bytes)
        agreed = getattr(self, "_comm_sizes", {})
        if key not in agreed:
            t = torch.tensor([key], dtype=torch.int64, device=self.data.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            agreed[key] = int(t.item())
            self._comm_sizes = agreed
        need = agreed[key]
        if comm is None or comm.heap_bytes < need:
            if comm is not None:
                comm.close()
            comm = PeerComm(self.group, need)
            self._comm = comm
        return comm

    def extended(self, ld, dtype=None):
        import torch

        return torch.zeros((self.plan.rows_alloc, ld), dtype=dtype or self.dtype, device=self.data.device)

    def middle(self, Xext):
        m0 = self.plan.row(self.r0)
        return Xext[m0:m0 + self.n]

    def matmat_extended(self, Xext, W):
        """``W[n_loc][ld] = A[r0:r1, :] @ X`` with `Xext` the extended block (halo filled)."""
        lib = _lib.load()
        st = self._struct()
        ld = Xext.shape[1]
        ws = getattr(self, "_mm_ws", None)
        if ws is None or ws[0] != ld:
            nbytes = lib.mf_matmat_workspace_bytes(ctypes.byref(st), ld)
            ws = (ld, _device.workspace(nbytes))
            self._mm_ws = ws
        _lib.check(lib.mf_matmat(ctypes.byref(st), Xext.data_ptr(), W.data_ptr(), ld,
                                 ws[1].data_ptr(), ws[1].numel(), _device.stream()))

    def matmat_blocked(self, X):
        import torch

        Xext = self.extended(X.shape[1], X.dtype)
        self.middle(Xext).copy_(X)
        exchange_halo(self.plan, Xext, self.group)
        W = torch.empty_like(X)
        self.matmat_extended(Xext, W)
        return W


# ----------------------------------------------------------------------------- backends


class CudaBackend:
    """The vector kernels of the row-sharded drivers, through the C ABI."""

    def __init__(self, ld, max_nq=4):
        self.lib = _lib.load()
        self.ld = ld
        self.ws = _device.workspace(self.lib.mf_blockvec_workspace_bytes(ld, max_nq))

    def _w(self):
        return self.ws.data_ptr(), self.ws.numel(), _device.stream()

    @staticmethod
    def empty(shape, like):
        import torch

        return torch.empty(shape, dtype=like.dtype, device=like.device)

    @staticmethod
    def sums(shape, like):
        import torch

        return torch.zeros(shape, dtype=torch.float64, device=like.device)

    def block_dot(self, X, Y, sums):
        n, ld = X.shape
        _lib.check(self.lib.mf_block_dot(X.data_ptr(), Y.data_ptr(), _device.mf_dtype(X.dtype), n, ld,
                                         sums.data_ptr(), *self._w()))

    def reorth_dots(self, Q, nq, V, sums):
        n, ld = V.shape
        _lib.check(self.lib.mf_reorth_dots(Q.data_ptr(), nq, V.data_ptr(), _device.mf_dtype(V.dtype), n,
                                           ld, sums.data_ptr(), *self._w()))

    def reorth_update(self, Q, nq, h, V, sqnorm=None):
        n, ld = V.shape
        _lib.check(self.lib.mf_reorth_update(Q.data_ptr(), nq, h.data_ptr(), V.data_ptr(),
                                             _device.mf_dtype(V.dtype), n, ld,
                                             None if sqnorm is None else sqnorm.data_ptr(), *self._w()))

    def lanczos_update(self, W, Rc, a, Rp, bprev, out, sqnorm):
        n, ld = W.shape
        _lib.check(self.lib.mf_lanczos_update(W.data_ptr(), Rc.data_ptr(), None, a.data_ptr(),
                                              None if Rp is None else Rp.data_ptr(), None,
                                              None if bprev is None else bprev.data_ptr(),
                                              out.data_ptr(), _device.mf_dtype(W.dtype), n, ld,
                                              sqnorm.data_ptr(), *self._w()))

    def scale(self, X, s, out, divide):
        n, ld = X.shape
        _lib.check(self.lib.mf_block_scale(X.data_ptr(), s.data_ptr(), out.data_ptr(), int(divide),
                                           _device.mf_dtype(X.dtype), n, ld, _device.stream()))

    def finalize(self, sums, take_sqrt, value=None, inv=None):
        like = value if value is not None else inv
        _lib.check(self.lib.mf_sums_finalize(sums.data_ptr(), sums.numel(), int(take_sqrt),
                                             None if value is None else value.data_ptr(),
                                             None if inv is None else inv.data_ptr(),
                                             _device.mf_dtype(like.dtype), _device.stream()))

    def full_offdiag(self, offdiag_row, h_row):
        _lib.check(self.lib.mf_full_offdiag(offdiag_row.data_ptr(), h_row.data_ptr(),
                                            _device.mf_dtype(h_row.dtype), h_row.numel(),
                                            _device.stream()))

    @staticmethod
    def matmat(op, Xext, W):
        op.matmat_extended(Xext, W)


# ----------------------------------------------------------------------------- drivers


def lanczos_sharded_native(op, V0, k, reortho, *, want_Q, want_residual):
    """`mf_lanczos_sharded`: the whole k-step loop of a row-sharded decomposition in one call --
    halos pushed into the neighbours' extended blocks and every reduction all-reduced over peer
    memory inside the reducing kernel (no NCCL, no host work between the kernels)."""
    import torch

    lib = _lib.load()
    nloc, ld = V0.shape
    dt, dev = V0.dtype, V0.device
    full = reortho == "full"
    rflag = _lib.MF_REORTHO_FULL if full else _lib.MF_REORTHO_NONE
    keep_Q = bool(want_Q or full)
    plan = op.plan.c_struct()
    st = op._struct()
    mfdt = _device.mf_dtype(dt)
    need = lib.mf_lanczos_sharded_heap_bytes(ctypes.byref(plan), ld, k, rflag, int(keep_Q), mfdt)
    comm = op.peer_comm(need)
    nblocks = max(k, 1) if keep_Q else 2
    if comm is None:
        ext = torch.empty((nblocks, op.plan.rows_alloc, ld), dtype=dt, device=dev)
    else:
        ext = comm.heap_view(0, (nblocks, op.plan.rows_alloc, ld), dt)
    ws_bytes = lib.mf_lanczos_sharded_workspace_bytes(ctypes.byref(st), ld, k, rflag)
    if ws_bytes < 0:
        _lib.check(-1)
    ws = _device.workspace(ws_bytes)
    alphas = torch.empty((max(k, 1), ld), dtype=dt, device=dev)
    betas = torch.empty((max(k, 1), ld), dtype=dt, device=dev)
    init_len = torch.empty((ld,), dtype=dt, device=dev)
    residual = torch.empty((nloc, ld), dtype=dt, device=dev) if want_residual else None
    _lib.check(lib.mf_lanczos_sharded(None if comm is None else comm.handle, ctypes.byref(st),
                                      ctypes.byref(plan), V0.data_ptr(), ld, k, rflag, int(keep_Q), 0,
                                      ext.data_ptr(), alphas.data_ptr(), betas.data_ptr(),
                                      init_len.data_ptr(),
                                      None if residual is None else residual.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _device.stream()))
    Q = None
    if keep_Q and k > 0:
        m0 = op.plan.row(op.r0)
        Q = ext[:k, m0:m0 + nloc].clone()  # a copy: the heap is reused by the next call
    return alphas[:k], betas[:k], init_len, Q, residual


def lanczos_full_sharded(op, V0, k, *, backend=None, want_residual=True):
    """Arnoldi with CGS twice and ``T = (H + H^T)/2`` on a row-sharded operator
    (`matfree/decomp.py:426-477,130-143`), for a block of `ld` start vectors `V0[n_loc][ld]`.

    Returns ``(alphas [k][ld], betas [k][ld], init_len [ld], Q [k][n_loc][ld], residual)`` with the
    same meaning as `mf_lanczos` (betas row k-1 = norm of the last residual)."""
    if backend is None and _use_peer_memory(op.group):
        return lanczos_sharded_native(op, V0, k, "full", want_Q=True, want_residual=want_residual)
    be = backend or CudaBackend(V0.shape[1], max_nq=k)
    group = op.group
    nloc, ld = V0.shape
    Q = be.empty((max(k, 1), nloc, ld), V0)
    V = be.empty((nloc, ld), V0)
    Xext = op.extended(ld, V0.dtype)
    alphas = be.empty((max(k, 1), ld), V0)
    betas = be.empty((max(k, 1), ld), V0)
    h = be.empty((max(k, 1), ld), V0)
    init_len = be.empty((ld,), V0)
    sums = be.sums((max(k, 1), ld), V0)
    sq = be.sums((ld,), V0)

    be.block_dot(V0, V0, sq)
    _all_reduce(sq, group)
    be.finalize(sq, True, value=init_len)
    cur, length = V0, init_len
    for i in range(k):
        be.scale(cur, length, Q[i], True)                       # decomp.py:456-457
        op.middle(Xext).copy_(Q[i])
        exchange_halo(op.plan, Xext, group)
        be.matmat(op, Xext, V)                                  # :460
        be.reorth_dots(Q, i + 1, V, sums[: i + 1])              # :463 (filled columns only)
        _all_reduce(sums[: i + 1], group)
        be.finalize(sums[: i + 1], False, value=h[: i + 1])
        alphas[i].copy_(h[i])
        if i > 0:
            be.full_offdiag(betas[i - 1], h[i - 1])             # T = (H + H^T)/2, :133-135
        be.reorth_update(Q, i + 1, h, V)                        # :464
        be.reorth_dots(Q, i + 1, V, sums[: i + 1])              # :468
        _all_reduce(sums[: i + 1], group)
        be.finalize(sums[: i + 1], False, value=h[: i + 1])
        be.reorth_update(Q, i + 1, h, V, sq)                    # :468, norm fused (:471)
        _all_reduce(sq, group)
        be.finalize(sq, True, value=betas[i])
        cur, length = V, betas[i]
    residual = None
    if want_residual:
        residual = V if k > 0 else V0.clone()
    return alphas[:k], betas[:k], init_len, Q[:k], residual


def lanczos_none_sharded(op, V0, k, *, backend=None, want_Q=False, want_residual=True):
    """Three-term Lanczos on a row-sharded operator (`matfree/decomp.py:220-292`, the operation
    order of the reference: normalise, matvec, alpha, update, beta)."""
    if backend is None and _use_peer_memory(op.group):
        return lanczos_sharded_native(op, V0, k, "none", want_Q=want_Q, want_residual=want_residual)
    be = backend or CudaBackend(V0.shape[1])
    group = op.group
    nloc, ld = V0.shape
    Xa, Xb = op.extended(ld, V0.dtype), op.extended(ld, V0.dtype)
    W = be.empty((nloc, ld), V0)
    R = be.empty((nloc, ld), V0)
    Q = be.empty((k, nloc, ld), V0) if (want_Q and k > 0) else None
    alphas = be.empty((max(k, 1), ld), V0)
    betas = be.empty((max(k, 1), ld), V0)
    init_len = be.empty((ld,), V0)
    sq = be.sums((ld,), V0)

    be.block_dot(V0, V0, sq)
    _all_reduce(sq, group)
    be.finalize(sq, True, value=init_len)
    cur, length = V0, init_len
    prev_mid = None
    for j in range(k):
        Xc = Xa if j % 2 == 0 else Xb
        mid = op.middle(Xc)
        be.scale(cur, length, mid, True)                        # v_j = r / b   (decomp.py:227,291)
        if Q is not None:
            Q[j].copy_(mid)
        exchange_halo(op.plan, Xc, group)
        be.matmat(op, Xc, W)                                    # :287
        be.block_dot(mid, W, sq)                                # :288
        _all_reduce(sq, group)
        be.finalize(sq, False, value=alphas[j])
        be.lanczos_update(W, mid, alphas[j], prev_mid, betas[j - 1] if j > 0 else None, R, sq)  # :289
        _all_reduce(sq, group)
        be.finalize(sq, True, value=betas[j])                   # :290
        cur, length, prev_mid = R, betas[j], mid
    residual = None
    if want_residual:
        residual = R if k > 0 else V0.clone()
    return alphas[:k], betas[:k], init_len, Q, residual

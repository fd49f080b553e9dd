"""Row-sharded drivers on CPU (`gloo`, world size 2 and 3): partitioning, halo plan, halo
exchange and the placement of the all-reduces in `matfree_b200/_rowshard.py`, checked against
the single-process oracle.  The vector kernels are replaced by a NumPy backend here (the CUDA
ones are covered by the `-m gpu` tests); the step loop, the plan and the collectives are the
product code."""

import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
import scipy.sparse as sp
sys.path.insert(0, os.environ["MF_ROOT"])
from matfree_b200 import _rowshard, workloads
from oracle import ref, prng as oprng

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shape = (7, 5, 4)
n = int(np.prod(shape)); plane = shape[1] * shape[2]
k, ld = 9, 4


class NumpyBackend:
    # same contract as _rowshard.CudaBackend, on CPU tensors
    @staticmethod
    def empty(shp, like): return torch.zeros(shp, dtype=like.dtype)
    @staticmethod
    def sums(shp, like): return torch.zeros(shp, dtype=torch.float64)
    def block_dot(self, X, Y, sums): sums.copy_((X.double() * Y.double()).sum(0))
    def reorth_dots(self, Q, nq, V, sums): sums.copy_(torch.einsum("jrc,rc->jc", Q[:nq].double(), V.double()))
    def reorth_update(self, Q, nq, h, V, sqnorm=None):
        V.sub_(torch.einsum("jrc,jc->rc", Q[:nq], h[:nq]))
        if sqnorm is not None: sqnorm.copy_((V.double() ** 2).sum(0))
    def lanczos_update(self, W, Rc, a, Rp, bprev, out, sqnorm):
        t = W - a * Rc
        if Rp is not None: t = t - bprev * Rp
        out.copy_(t); sqnorm.copy_((t.double() ** 2).sum(0))
    def scale(self, X, s, out, divide): out.copy_(X / s if divide else X * s)
    def finalize(self, sums, take_sqrt, value=None, inv=None):
        v = torch.sqrt(sums) if take_sqrt else sums
        if value is not None: value.copy_(v.reshape(value.shape).to(value.dtype))
        if inv is not None: inv.copy_((1.0 / v).reshape(inv.shape).to(inv.dtype))
    def full_offdiag(self, off, h): off.copy_(0.5 * (off + h))
    @staticmethod
    def matmat(op, Xext, W): W.copy_(torch.from_numpy(op.A_local @ Xext.numpy()))


class CpuShardedOp:
    # the parts of RowShardedCsr the drivers use, without a GPU
    def __init__(self, ip, ix, d, n, r0, dtype):
        self.group = None
        self.n = ip.numel() - 1
        self.r0, self.r1 = r0, r0 + self.n
        cmin, cmax = int(ix.min()), int(ix.max()) + 1
        self.plan = _rowshard.make_plan(self.r0, self.r1, min(cmin, self.r0), max(cmax, self.r1), None)
        self.A_local = sp.csr_matrix((d.numpy(), (ix + (self.plan.pad - self.plan.c0)).numpy(), ip.numpy()),
                                     shape=(self.n, self.plan.rows_alloc))
        self.dtype = dtype
    def extended(self, ld, dtype=None): return torch.zeros((self.plan.rows_alloc, ld), dtype=dtype or self.dtype)
    def middle(self, X): return X[self.plan.row(self.r0):self.plan.row(self.r0) + self.n]


for mode in ("stencil", "unbanded"):
    if mode == "stencil":
        r0, r1 = _rowshard.slab_range(n, world, rank, align=plane)
        ip, ix, d = workloads.laplacian_csr_rows(shape, r0, r1, shift=1.0, dtype="float64")
        ipf, ixf, df = workloads.laplacian_csr(shape, shift=1.0, dtype="float64")
        A = sp.csr_matrix((df.numpy(), ixf.numpy(), ipf.numpy()), shape=(n, n))
    else:
        # symmetric matrix without band structure: every rank needs (almost) the whole vector
        rng = np.random.default_rng(3)
        B = sp.random(n, n, density=0.05, random_state=rng, format="csr")
        A = (B + B.T + sp.identity(n) * 8.0).tocsr()
        A.sort_indices()
        r0, r1 = _rowshard.slab_range(n, world, rank)
        Al = A[r0:r1]
        ip, ix, d = (torch.from_numpy(Al.indptr.astype(np.int32)), torch.from_numpy(Al.indices.astype(np.int64)),
                     torch.from_numpy(Al.data))
    op = CpuShardedOp(ip, ix, d, n, r0, torch.float64)
    if mode == "stencil" and world > 1:
        # one plane of halo per interior side, exchanged with the direct neighbours only
        assert op.plan.lo == (plane if rank > 0 else 0) and op.plan.hi == (plane if rank < world - 1 else 0)
        assert all(abs(p - rank) == 1 for p, _, _ in op.plan.recvs + op.plan.sends)
    V = oprng.normal(oprng.prng_key(11), (ld, n), np.float64)      # (P, n)
    V0 = torch.from_numpy(np.ascontiguousarray(V.T[r0:r1]))         # my slab, blocked [n_loc][ld]
    be = NumpyBackend()
    matmat = lambda X: (A @ X.T).T

    a, b, ln, Q, res = _rowshard.lanczos_full_sharded(op, V0, k, backend=be)
    od, oe, ol = ref.lanczos_full_batched(matmat, V, k)
    assert np.allclose(a.numpy().T, od, rtol=1e-10, atol=1e-12), mode
    assert np.allclose(b.numpy().T[:, : k - 1], oe, rtol=1e-10, atol=1e-12), mode
    assert np.allclose(ln.numpy(), ol, rtol=1e-13)
    # the local basis rows are orthonormal once summed over the ranks
    G = torch.einsum("jrc,lrc->cjl", Q.double(), Q.double())
    dist.all_reduce(G)
    assert np.allclose(G.numpy(), np.broadcast_to(np.eye(k), (ld, k, k)), atol=1e-10)

    a, b, ln, Q, res = _rowshard.lanczos_none_sharded(op, V0, k, backend=be, want_Q=True)
    oa, ob, ol = ref.lanczos_none_batched(matmat, V, k)
    assert np.allclose(a.numpy().T, oa, rtol=1e-9, atol=1e-11), mode
    assert np.allclose(b.numpy().T, ob, rtol=1e-9, atol=1e-11), mode
    # residual = b_{k-1} v_k  (decomp.py:167): orthogonal to v_{k-1}
    dot = (res.double() * Q[k - 1].double()).sum(0)
    dist.all_reduce(dot)
    assert np.all(np.abs(dot.numpy()) < 1e-8)

dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_halo_plan_is_consistent():
    from matfree_b200 import _rowshard

    n, world, plane = 120, 4, 20
    ranges = [_rowshard.slab_range(n, world, r, align=plane) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    needs = [(max(0, r0 - plane), min(n, r1 + plane)) for r0, r1 in ranges]
    plans = [_rowshard.HaloPlan(r, ranges, needs) for r in range(world)]
    for r, p in enumerate(plans):
        # every receive has the matching send on the peer, with the same row range
        for peer, a, b in p.recvs:
            assert (r, a, b) in plans[peer].sends
        for peer, a, b in p.sends:
            assert (r, a, b) in plans[peer].recvs
        assert p.n_ext == p.lo + (p.r1 - p.r0) + p.hi == p.halo_rows + (p.r1 - p.r0)
    # a rank whose rows touch every column needs everything that is not its own
    wide = _rowshard.HaloPlan(1, ranges, [needs[0], (0, n), needs[2], needs[3]])
    assert wide.halo_rows == n - (ranges[1][1] - ranges[1][0])
    # ragged: more ranks than planes leaves empty slabs at the end
    rag = [_rowshard.slab_range(40, 4, r, align=20) for r in range(4)]
    assert rag == [(0, 20), (20, 40), (40, 40), (40, 40)]


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_lanczos_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MF_ROOT=ROOT, OMP_NUM_THREADS="2")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), str(script)],
        env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world

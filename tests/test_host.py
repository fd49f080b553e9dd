"""CPU tests: the C-ABI library loads and exports every declared symbol; host logic."""

import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import prng as oprng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from matfree_b200 import _lib

    lib = _lib.load()  # raises if missing or a symbol is absent
    header = open(os.path.join(ROOT, "include", "matfree_b200.h")).read()
    declared = set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", header))
    declared -= {"mf_operator"}
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.mf_abi_version() == 7
    assert lib.mf_launch_count() == 0


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call (no GPU needed)."""
    import ctypes

    from matfree_b200 import _lib

    lib = _lib.load()
    op = _lib.MfOperator(kind=_lib.MF_OP_CSR, dtype=0, n=13, m=13, nnz=1, values=8, indptr=8, indices=8,
                         lda=0, split_planes=None)
    assert lib.mf_lanczos_workspace_bytes(ctypes.byref(op), 16, 14, 0, 0) == -1
    assert b"exceeds the acceptable range" in lib.mf_last_error()
    assert lib.mf_lanczos_workspace_bytes(ctypes.byref(op), 16, -1, 0, 0) == -1
    assert lib.mf_lanczos_workspace_bytes(ctypes.byref(op), 24, 3, 0, 0) == -1  # ld not a power of two
    assert lib.mf_lanczos_workspace_bytes(ctypes.byref(op), 16, 3, 0, 0) > 0
    assert lib.mf_estimate_workspace_bytes(ctypes.byref(op), 16, 3, 1, 0) > 0
    with pytest.raises(ValueError, match="exceeds"):
        _lib.check(lib.mf_lanczos(ctypes.byref(op), 8, 16, 14, 0, 8, 8, 8, None, None, 8, 1 << 20, None))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from matfree_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryMissingError, match="no CPU fallback"):
        _lib.load()


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import matfree_b200 as m

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.ops.dense(np.eye(3, dtype=np.float32))


def test_product_does_not_import_oracle():
    out = subprocess.run(
        [sys.executable, "-c",
         "import sys, matfree_b200, matfree_b200.workloads; "
         "print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"],
        cwd=ROOT, capture_output=True, text=True, check=True)
    assert out.stdout.strip() == "False"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "matfree_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_prng_key_and_split_match_oracle():
    from matfree_b200.backend import prng

    for seed in (0, 1, 42, 2**40 + 7):
        assert np.array_equal(prng.prng_key(seed), oprng.prng_key(seed))
        assert np.array_equal(prng.split(prng.prng_key(seed), 5), oprng.split(oprng.prng_key(seed), 5))


def test_factories_validate_like_the_reference():
    import matfree_b200 as m

    with pytest.raises(ValueError, match="unsupported"):
        m.decomp.tridiag_sym(3, reortho="partial")
    tri = m.decomp.tridiag_sym(4, reortho="none", materialize=False)
    assert tri._mf_spec == {"kind": "tridiag_sym", "num_matvecs": 4, "reortho": "none", "materialize": False}
    integ = m.funm.integrand_funm_sym_logdet(tri)
    assert integ._mf_integrand["kind"] == "slq" and integ._mf_integrand["matfun"] is np.log
    assert m.stochtrace.monte_carlo_trace()._mf_integrand == {"kind": "trace"}


def test_shard_range_partitions_probes():
    from matfree_b200._sharding import shard_range

    for P in (1, 7, 8, 1000, 8192):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(P, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == P
            for (a0, a1), (b0, b1) in zip(got[:-1], got[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(b - a for a, b in got) == P


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MF_ROOT"])
from matfree_b200._sharding import shard_range, gather_shards
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P = int(os.environ["MF_P"])
full = torch.arange(P, dtype=torch.float32) * 0.5 + 1.0
p0, p1 = shard_range(P, world, rank)
got = gather_shards(full[p0:p1].clone(), P)
assert torch.equal(got, full), (rank, got, full)
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("P", [8, 13])
def test_gather_shards_world_size_2_gloo(tmp_path, P):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MF_ROOT=ROOT, MF_P=str(P), MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(29531 + P), str(script)],
        env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


@pytest.mark.parametrize("world,n,band", [(2, 64, 5), (3, 300, 10), (4, 257, 3), (8, 4096, 256), (3, 40, 39)])
def test_peer_halo_plans_are_mutually_consistent(world, n, band):
    """`HaloPlan.c_struct()` (the `mf_halo_plan_t` the peer-memory halo kernel executes): replaying
    every rank's send list with NumPy must leave every extended block equal to the global vector
    on exactly the rows the rank's columns reference -- i.e. each send lands where the RECEIVER's
    plan expects it -- and `recv_peers` must be exactly the ranks that send to it."""
    from matfree_b200 import _rowshard

    ranges = [_rowshard.slab_range(n, world, r, align=1) for r in range(world)]
    needs = [(max(0, a - band), min(n, b + band)) for a, b in ranges]
    plans = [_rowshard.HaloPlan(r, ranges, needs) for r in range(world)]
    x = np.arange(1, n + 1, dtype=np.float64)
    ext = []
    for p in plans:
        c = p.c_struct()
        assert c.rows_alloc % 4 == 0 and c.mid_row % 4 == 0 and c.mid_row + (p.r1 - p.r0) <= c.rows_alloc
        e = np.zeros(c.rows_alloc)
        e[c.mid_row:c.mid_row + (p.r1 - p.r0)] = x[p.r0:p.r1]
        ext.append(e)
    senders = [set() for _ in range(world)]
    for r, p in enumerate(plans):
        c = p.c_struct()
        for i in range(c.num_sends):
            s = c.sends[i]
            assert s.dst_rows_alloc == plans[s.peer].rows_alloc
            ext[s.peer][s.dst_row:s.dst_row + s.rows] = ext[r][s.src_row:s.src_row + s.rows]
            senders[s.peer].add(r)
    for r, p in enumerate(plans):
        c = p.c_struct()
        assert sorted(senders[r]) == [c.recv_peers[i] for i in range(c.num_recv_peers)]
        lo, hi = p.row(p.c0), p.row(p.c1)
        assert np.array_equal(ext[r][lo:hi], x[p.c0:p.c1]), r


def test_sharded_entry_points_validate_without_gpu():
    import ctypes

    from matfree_b200 import _lib, _rowshard

    lib = _lib.load()
    plan = _rowshard.HaloPlan(0, [(0, 13)], [(0, 13)]).c_struct()
    op = _lib.MfOperator(kind=_lib.MF_OP_CSR, dtype=0, n=13, m=13, nnz=1, values=8, indptr=8, indices=8,
                         lda=0, split_planes=None)
    assert lib.mf_lanczos_sharded_heap_bytes(ctypes.byref(plan), 4, 10, _lib.MF_REORTHO_FULL, 0, 0) == 10 * 16 * 4 * 4
    assert lib.mf_lanczos_sharded_heap_bytes(ctypes.byref(plan), 4, 10, _lib.MF_REORTHO_NONE, 0, 0) == 2 * 16 * 4 * 4
    assert lib.mf_lanczos_sharded_workspace_bytes(ctypes.byref(op), 4, 10, _lib.MF_REORTHO_FULL) > 0
    assert lib.mf_lanczos_sharded_workspace_bytes(ctypes.byref(op), 3, 10, _lib.MF_REORTHO_FULL) == -1
    dense = _lib.MfOperator(kind=_lib.MF_OP_DENSE, dtype=0, n=13, m=13, nnz=0, values=8, indptr=None,
                            indices=None, lda=13, split_planes=None)
    assert lib.mf_lanczos_sharded_workspace_bytes(ctypes.byref(dense), 4, 10, 0) == -1  # CSR only
    h = ctypes.c_void_p()
    with pytest.raises(ValueError, match="world must be"):
        _lib.check(lib.mf_comm_create(9, 0, 0, ctypes.byref(h)))


def test_ravel_pytree_matches_jax_flattening_order():
    """`matfree/backend/tree.py:15-20` (jax.flatten_util.ravel_pytree): dict keys in sorted order,
    sequences in order, leaves raveled row-major; unravel and its batched form invert it."""
    import torch

    from matfree_b200.backend import tree

    t = {"b": np.arange(6.0).reshape(2, 3), "a": [(np.array([10.0, 11.0]),), np.float64(7.0)]}
    flat, unravel = tree.ravel_pytree(t, device="cpu")
    assert flat.tolist() == [10.0, 11.0, 7.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    back = unravel(flat)
    assert list(back) == ["b", "a"] and tuple(back["b"].shape) == (2, 3) and back["a"][1].shape == ()
    assert torch.equal(back["a"][0][0], torch.tensor([10.0, 11.0], dtype=torch.float64))
    rows = torch.stack([flat, 2 * flat])
    bb = unravel.batched(rows)
    assert tuple(bb["b"].shape) == (2, 2, 3) and torch.equal(bb["b"][1], 2 * back["b"])
    # the reference tests' vector pytree `[(v,)]` (tests/test_decomp/test_tridiag_sym.py:19-22)
    flat2, unravel2 = tree.ravel_pytree([(np.ones(4, np.float32),)], device="cpu")
    [(x,)] = unravel2(flat2)
    assert tuple(x.shape) == (4,) and flat2.dtype == torch.float32
    [(xb,)] = unravel2.batched(torch.zeros(3, 4))
    assert tuple(xb.shape) == (3, 4)
    # a flat array is its own pytree
    flat3, unravel3 = tree.ravel_pytree(np.ones(5, np.float32), device="cpu")
    assert unravel3.trivial and unravel3(flat3) is flat3


def test_ffi_shim_compiles_and_matches_integration_doc(tmp_path):
    """The XLA-FFI shim (csrc/ffi_xla.cc) is compiled against a stand-in of the FFI API
    (tests/ffi_stub): every handler's signature must match its binding (static_assert in `To`),
    every `*_ffi` target INTEGRATION.md registers must be defined by the shim, and every mf_*
    entry point the shim calls must be exported by the library with the header's signature
    (the compiler checks the call against include/matfree_b200.h)."""
    from matfree_b200 import _lib

    shim = os.path.join(ROOT, "matfree_b200", "csrc", "ffi_xla.cc")
    obj = str(tmp_path / "ffi_xla.o")
    cuda_inc = "/usr/local/cuda/include"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "tests", "ffi_stub"),
                        "-I", cuda_inc, shim, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    nm = subprocess.run(["nm", obj], capture_output=True, text=True, check=True).stdout
    defined = set(re.findall(r" T (mf_[a-z0-9_]+_ffi)\b", nm))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    in_doc = set(re.findall(r"\b(mf_[a-z0-9_]+_ffi)\b", doc))
    # brace patterns like mf_lanczos_{csr,dense,gram}_ffi are prose; the registration loop spells all names
    assert in_doc and in_doc <= defined, sorted(in_doc - defined)
    assert len(defined) == 10
    registered = set(re.findall(r'"(mf_[a-z0-9_]+_ffi)"', doc))
    assert registered == defined, sorted(registered ^ defined)
    # every ffi_call target in the doc is a registered name without the suffix
    targets = set(re.findall(r'ffi_call\("(mf_[a-z0-9_]+)"', doc))
    assert targets and all(t + "_ffi" in defined for t in targets), sorted(targets)
    src = open(shim).read()
    called = set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", src)) - {"mf_dtype_of"}
    called = {c for c in called if not c.endswith("_ffi")}
    lib = _lib.load()
    for name in called:
        assert hasattr(lib, name) and name in _lib.SIGNATURES, name


def test_integration_doc_csr_hints_on_known_stencils():
    """The `csr_hints` helper INTEGRATION.md gives the reference-side binding (NumPy, run once per
    operator) must produce the structure hints of include/matfree_b200.h that `matfree_b200.ops.csr`
    computes on the device: bandwidth, number of diagonals, line stride of 2-D / 3-D grid stencils,
    and "not a stencil" for an irregular matrix."""
    import re

    import numpy as np
    import scipy.sparse as sp

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"^def csr_hints\(.*?(?=^def )", doc, re.S | re.M)
    assert m, "INTEGRATION.md no longer defines csr_hints"
    ns = {"np": np}
    exec(m.group(0), ns)
    csr_hints = ns["csr_hints"]

    def lap(shape):
        mats = [sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(k, k)) for k in shape]
        out = None
        for i, T in enumerate(mats):
            term = None
            for j, k in enumerate(shape):
                f = T if i == j else sp.identity(k)
                term = f if term is None else sp.kron(term, f)
            out = term if out is None else out + term
        A = sp.csr_matrix(out)
        A.sort_indices()
        return A

    A2 = lap((6, 32))
    h = csr_hints(A2.indptr, A2.indices)
    assert (int(h["csr_bandwidth"]), int(h["csr_num_diagonals"]), int(h["csr_line_stride"])) == (32, 5, 32)
    A3 = lap((3, 4, 16))
    h = csr_hints(A3.indptr, A3.indices)
    assert (int(h["csr_bandwidth"]), int(h["csr_num_diagonals"]), int(h["csr_line_stride"])) == (64, 7, 16)
    B = sp.random(200, 200, density=0.02, random_state=np.random.default_rng(0), format="csr")
    B.sort_indices()
    h = csr_hints(B.indptr, B.indices)
    assert int(h["csr_num_diagonals"]) == 255 and int(h["csr_line_stride"]) == 0


def test_blocked_row_order_is_a_permutation_of_the_chunks(tmp_path):
    """`chunk_row0` / `choose_row_order` (csrc/spmm_common.cuh: the blocked row order of the CSR product
    for 3-D stencils) checked on the host: every chunk start is visited exactly once and consecutive
    blocks are the same row range of consecutive planes (tests/csrc/row_order_check.cu, compiled with
    nvcc; no CUDA call is made, so no GPU is needed)."""
    import shutil

    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "row_order_check")
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
                        "-I", os.path.join(ROOT, "matfree_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                        "-o", exe, os.path.join(ROOT, "tests", "csrc", "row_order_check.cu")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "row order: ok" in r.stdout, r.stdout + r.stderr

"""ctypes binding of libmatfree_b200.so (the C ABI in include/matfree_b200.h)."""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int32, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MF_LIB_PATH") or os.path.join(_HERE, "_lib", "libmatfree_b200.so")

MF_F32, MF_F64 = 0, 1
MF_SAMPLER_SIGNS, MF_SAMPLER_NORMAL = 0, 1
MF_PRNG_X64_BITS = 1
MF_LAYOUT_PROBE_MAJOR, MF_LAYOUT_BLOCKED = 0, 1
MF_OP_DENSE, MF_OP_CSR, MF_OP_GRAM = 0, 1, 2
MF_REORTHO_NONE, MF_REORTHO_FULL = 0, 1
MF_FN_NONE, MF_FN_LOG, MF_FN_EXP, MF_FN_INV, MF_FN_SQRT, MF_FN_POW, MF_FN_IDENTITY, MF_FN_SIN = range(8)
MF_INTEGRAND_SLQ, MF_INTEGRAND_TRACE = 0, 1
KERNEL_CLASSES = ["probe_gen", "spmm_csr", "gemm", "dot", "finalize", "lanczos_update", "scale",
                  "reorth_dots", "reorth_update", "tridiag_quad", "mc_reduce", "other"]


class MfOperator(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("dtype", c_int32),
        ("n", c_int64),
        ("m", c_int64),
        ("nnz", c_int64),
        ("values", c_void_p),
        ("indptr", c_void_p),
        ("indices", c_void_p),
        ("lda", c_int64),
        ("split_planes", c_void_p),
        ("csr_max_row_nnz", c_int32),
        ("csr_bandwidth", c_int64),
        ("csr_num_diagonals", c_int32),
        ("csr_line_stride", c_int64),
    ]


class MfHaloSend(Structure):
    _fields_ = [("peer", c_int32), ("src_row", c_int64), ("rows", c_int64), ("dst_row", c_int64),
                ("dst_rows_alloc", c_int64)]


class MfHaloPlan(Structure):
    _fields_ = [("rows_alloc", c_int64), ("mid_row", c_int64), ("num_sends", c_int32),
                ("sends", POINTER(MfHaloSend)), ("num_recv_peers", c_int32),
                ("recv_peers", POINTER(c_int32))]


MF_COMM_HANDLE_BYTES = 64

# name -> (restype, argtypes); every symbol include/matfree_b200.h declares
_OP = POINTER(MfOperator)
SIGNATURES = {
    "mf_last_error": (c_char_p, []),
    "mf_abi_version": (c_int32, []),
    "mf_launch_count": (c_int64, []),
    "mf_timing_enable": (c_int32, [c_int32]),
    "mf_timing_collect": (c_int32, [c_void_p, c_void_p]),
    "mf_probe_gen": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_int64, c_int64, c_int64,
                               c_uint32, c_uint32, c_int32, c_int32, c_void_p, c_void_p]),
    "mf_probe_gen_rows": (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_int64, c_int64, c_int64,
                                    c_int64, c_uint32, c_uint32, c_int32, c_int32, c_void_p]),
    "mf_gemm_config": (c_int32, [c_int32, c_int32]),
    "mf_spmm_config": (c_int32, [c_int32, c_int32, c_int32, c_int32]),
    "mf_operator_split_bytes": (c_int64, [_OP]),
    "mf_operator_split": (c_int32, [_OP, c_void_p, c_void_p]),
    "mf_matmat_workspace_bytes": (c_int64, [_OP, c_int64]),
    "mf_matmat": (c_int32, [_OP, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mf_matmat_dense": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p,
                                  c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mf_matmat_csr": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_matmat_gram": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mf_to_blocked": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p]),
    "mf_from_blocked": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p]),
    "mf_lanczos_workspace_bytes": (c_int64, [_OP, c_int64, c_int64, c_int32, c_int32]),
    "mf_lanczos": (c_int32, [_OP, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_hessenberg_workspace_bytes": (c_int64, [_OP, c_int64, c_int64]),
    "mf_hessenberg": (c_int32, [_OP, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_blockvec_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "mf_block_dot": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                               c_int64, c_void_p]),
    "mf_reorth_dots": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_int64, c_int64, c_void_p,
                                 c_void_p, c_int64, c_void_p]),
    "mf_reorth_update": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int64, c_int64,
                                   c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_lanczos_update": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p,
                                    c_void_p, c_int64, c_void_p]),
    "mf_block_scale": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64,
                                 c_void_p]),
    "mf_sums_finalize": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32,
                                   c_void_p]),
    "mf_full_offdiag": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_void_p]),
    "mf_tridiag_quad_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "mf_tridiag_quad": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64,
                                  c_int32, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_void_p]),
    "mf_mc_reduce": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "mf_estimate_workspace_bytes": (c_int64, [_OP, c_int64, c_int64, c_int32, c_int32]),
    "mf_estimate": (c_int32, [_OP, c_int32, c_int32, c_int32, c_uint32, c_uint32, c_int64, c_int64,
                              c_int64, c_int64, c_int32, c_int32, c_double, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_slq_estimate_dense": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32,
                                        c_uint32, c_uint32, c_int64, c_int64, c_int64, c_int64,
                                        c_int32, c_int32, c_double, c_void_p, c_void_p, c_int64,
                                        c_void_p]),
    "mf_slq_estimate_csr": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                      c_int32, c_int32, c_uint32, c_uint32, c_int64, c_int64,
                                      c_int64, c_int64, c_int32, c_int32, c_double, c_void_p,
                                      c_void_p, c_int64, c_void_p]),
    "mf_slq_estimate_gram": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32,
                                       c_int32, c_uint32, c_uint32, c_int64, c_int64, c_int64,
                                       c_int64, c_int32, c_int32, c_double, c_void_p, c_void_p,
                                       c_int64, c_void_p]),
    "mf_funm_lanczos_workspace_bytes": (c_int64, [_OP, c_int64, c_int64]),
    "mf_funm_lanczos": (c_int32, [_OP, c_void_p, c_int64, c_int64, c_int64, c_int32, c_double,
                                  c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_tridiag_funm_e1": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64,
                                     c_int32, c_double, c_void_p, c_void_p, c_int64, c_void_p]),
    "mf_basis_combine": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64,
                                   c_int64, c_void_p, c_void_p]),
    "mf_matmat_rect_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int32, c_int32, c_int32,
                                                 c_int64]),
    "mf_matmat_rect": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mf_bidiag_quad": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64,
                                 c_int32, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p]),
    "mf_hutch_rows": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_int32,
                                c_void_p, c_void_p, c_void_p]),
    "mf_lincomb": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int64,
                             c_void_p]),
    "mf_sddmm_csr": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                               c_void_p, c_int32, c_void_p]),
    "mf_comm_create": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_void_p)]),
    "mf_comm_handle": (c_int32, [c_void_p, c_void_p]),
    "mf_comm_connect": (c_int32, [c_void_p, c_void_p]),
    "mf_comm_heap": (c_void_p, [c_void_p]),
    "mf_comm_heap_bytes": (c_int64, [c_void_p]),
    "mf_comm_status": (c_int32, [c_void_p, c_void_p]),
    "mf_comm_barrier": (c_int32, [c_void_p, c_void_p]),
    "mf_comm_disconnect": (c_int32, [c_void_p]),
    "mf_comm_destroy": (c_int32, [c_void_p]),
    "mf_halo_exchange": (c_int32, [c_void_p, POINTER(MfHaloPlan), c_int64, c_int64, c_int64,
                                   c_int32, c_int32, c_void_p]),
    "mf_lanczos_sharded_heap_bytes": (c_int64, [POINTER(MfHaloPlan), c_int64, c_int64, c_int32,
                                                c_int32, c_int32]),
    "mf_lanczos_sharded_workspace_bytes": (c_int64, [_OP, c_int64, c_int64, c_int32]),
    "mf_lanczos_sharded": (c_int32, [c_void_p, _OP, POINTER(MfHaloPlan), c_void_p, c_int64, c_int64,
                                     c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
}

_lib = None


class LibraryMissingError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissingError(
            f"{LIB_PATH} not found: matfree_b200 has no CPU fallback. Build the CUDA library "
            "first (`make -C matfree_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def timing_enable(on: bool):
    load().mf_timing_enable(int(bool(on)))


def timing_collect():
    """{kernel class: (milliseconds, launches)} since the last collect (synchronises)."""
    n = len(KERNEL_CLASSES)
    ms = (c_double * n)()
    cnt = (c_int64 * n)()
    load().mf_timing_collect(ms, cnt)
    return {KERNEL_CLASSES[i]: (ms[i], cnt[i]) for i in range(n) if cnt[i]}


class MatfreeError(RuntimeError):
    pass


def check(rc: int):
    """Turn a negative status into an exception (ValueError for invalid arguments,
    matching the exceptions matfree raises at trace time)."""
    if rc == 0:
        return
    msg = load().mf_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    raise MatfreeError(f"libmatfree_b200 error {rc}: {msg}")

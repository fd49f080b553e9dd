"""ctypes wrapper of `oracle/slq_port.c` (the multi-threaded C port).  TEST INFRASTRUCTURE."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libslq_port.so")
_lib = None


def build(force: bool = False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "slq_port.c")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = ctypes.CDLL(LIB)
        i64, u32, vp = ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p
        lib.slq_port_num_threads.restype = ctypes.c_int
        lib.slq_port_set_threads.restype = None
        lib.slq_port_set_threads.argtypes = [ctypes.c_int]
        lib.slq_port_csr_logdet.restype = ctypes.c_int
        lib.slq_port_csr_logdet.argtypes = [vp, vp, vp, i64, i64, i64, i64, u32, u32, vp, vp, vp]
        lib.slq_port_csr_trace.restype = ctypes.c_int
        lib.slq_port_csr_trace.argtypes = [vp, vp, vp, i64, i64, i64, u32, u32, vp]
        lib.slq_port_probes.restype = None
        lib.slq_port_probes.argtypes = [vp, i64, i64, i64, u32, u32]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(load().slq_port_num_threads())


def set_threads(n: int | None = None) -> int:
    """Use `n` OpenMP threads (default: every core this process may run on), whatever
    OMP_NUM_THREADS says -- `torch.distributed.run` sets it to 1 in its workers."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    load().slq_port_set_threads(int(n))
    return num_threads()


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def csr_logdet_quadforms(indptr, indices, data, key, p0, num, k, return_coeffs=False):
    """Per-probe SLQ log-det quadratic forms of probes p0..p0+num-1 (fp32, reortho none)."""
    lib = load()
    indptr = np.ascontiguousarray(indptr, np.int32)
    indices = np.ascontiguousarray(indices, np.int32)
    data = np.ascontiguousarray(data, np.float32)
    n = indptr.shape[0] - 1
    quad = np.empty(num, np.float32)
    al = np.empty((num, k), np.float32)
    be = np.empty((num, k), np.float32)
    rc = lib.slq_port_csr_logdet(_ptr(indptr), _ptr(indices), _ptr(data), n, num, p0, k, int(key[0]),
                                 int(key[1]), _ptr(quad), _ptr(al), _ptr(be))
    if rc != 0:
        raise MemoryError("slq_port_csr_logdet failed")
    return (quad, al, be) if return_coeffs else quad


def csr_trace_samples(indptr, indices, data, key, p0, num):
    lib = load()
    indptr = np.ascontiguousarray(indptr, np.int32)
    indices = np.ascontiguousarray(indices, np.int32)
    data = np.ascontiguousarray(data, np.float32)
    n = indptr.shape[0] - 1
    out = np.empty(num, np.float32)
    if lib.slq_port_csr_trace(_ptr(indptr), _ptr(indices), _ptr(data), n, num, p0, int(key[0]),
                              int(key[1]), _ptr(out)) != 0:
        raise MemoryError("slq_port_csr_trace failed")
    return out


def probes(n, num, p0, key):
    X = np.empty((n, num), np.float32)
    load().slq_port_probes(_ptr(X), n, num, p0, int(key[0]), int(key[1]))
    return X

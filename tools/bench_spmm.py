"""Time the CSR SpMM on the C2 operator (2-D 5-point Laplacian 4096^2, ld = 256): ms per launch and
achieved GB/s against the algorithmic bytes (read X once, write W, matrix once).  Tuning knobs come
from the environment (MF_SPMM_*), so one process = one configuration.

usage: python tools/bench_spmm.py [grid [ld [reps]]]
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import matfree_b200 as m  # noqa: E402
from matfree_b200 import _device, _lib, workloads  # noqa: E402


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ld = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    lib = _lib.load()
    n = g * g
    ip, ix, d = workloads.laplacian_csr((g, g), shift=1.0, device="cuda")
    op = m.ops.csr(ip, ix, d)
    X = torch.randn(n, ld, device="cuda")
    W = torch.empty_like(X)
    st = op._struct()
    ws = _device.workspace(lib.mf_matmat_workspace_bytes(ctypes.byref(st), ld))

    def run():
        _lib.check(lib.mf_matmat(ctypes.byref(st), X.data_ptr(), W.data_ptr(), ld, ws.data_ptr(), ws.numel(),
                                 _device.stream()))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 2 * n * ld * 4 + op.nnz * 8 + 4 * (n + 1)
    # spot check against a dense stencil evaluation of a few rows
    r = torch.tensor([0, 1, g, n // 2 + 7, n - 1], device="cuda")
    want = 5.0 * X[r]
    for off in (-1, 1, -g, g):
        c = r + off
        ok = (c >= 0) & (c < n)
        if abs(off) == 1:
            ok &= (c // g) == (r // g)
        want[ok] -= X[c[ok]]
    err = float((W[r] - want).abs().max())
    knobs = {k: v for k, v in os.environ.items() if k.startswith("MF_SPMM_")}
    print(json.dumps({"grid": g, "ld": ld, "ms": ms, "gbs": alg / ms / 1e6, "knobs": knobs, "spot_err": err}),
          flush=True)


if __name__ == "__main__":
    main()

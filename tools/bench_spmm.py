"""Time the CSR product on a stencil operator (default: BASELINE config 2, the 2-D 5-point Laplacian
4096^2, ld = 256): ms per launch and achieved GB/s against the algorithmic bytes (read X once,
write W, matrix once), for a list of kernel configurations in ONE process (`mf_spmm_config`).

usage: python tools/bench_spmm.py [--shape 4096,4096] [--ld 256] [--reps 10] [--dtype float32]
                                  [--configs band:rows:pfd:minb,...]
A checksum of W (sum of its bit patterns) is printed per configuration: all must agree.
"""
import argparse
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
import matfree_b200 as m  # noqa: E402
from matfree_b200 import _device, _lib, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4096,4096")
    ap.add_argument("--ld", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--configs", default="0:64:2:3,1:64:0:3,2:64:0:3")
    a = ap.parse_args()
    shape = tuple(int(x) for x in a.shape.split(","))
    ld = a.ld
    lib = _lib.load()
    n = 1
    for s in shape:
        n *= s
    ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, dtype=a.dtype, device="cuda")
    op = m.ops.csr(ip, ix, d)
    X = torch.randn(n, ld, device="cuda", dtype=d.dtype)
    W = torch.empty_like(X)
    st = op._struct()
    ws = _device.workspace(lib.mf_matmat_workspace_bytes(ctypes.byref(st), ld))
    es = X.element_size()
    alg = 2 * n * ld * es + op.nnz * (es + 4) + 4 * (n + 1)

    def run():
        _lib.check(lib.mf_matmat(ctypes.byref(st), X.data_ptr(), W.data_ptr(), ld, ws.data_ptr(), ws.numel(),
                                 _device.stream()))

    for cfg in a.configs.split(","):
        band, rows, pfd, minb = (int(x) for x in cfg.split(":"))
        lib.mf_spmm_config(band, rows, pfd, minb)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        view = W.view(torch.int32) if es == 4 else W.view(torch.int64)
        chk = int(view.to(torch.int64).sum())
        print(json.dumps({"shape": shape, "ld": ld, "dtype": a.dtype, "band": band, "rows": rows, "pfd": pfd,
                          "minb": minb, "ms": round(ms, 4), "gbs": round(alg / ms / 1e6, 1), "checksum": chk}),
              flush=True)


if __name__ == "__main__":
    main()

"""Time the dense / Gram probe-block contraction (tcgen05 3xTF32 vs CUDA cores).

usage: python tools/bench_gemm.py [m n ld [reps]]   (defaults: C3 shape 65536 16384 256)
Prints one JSON line per (variant, kernel): ms per Gram matmat and fp32-equivalent TFLOP/s
(4 m n ld flops per Gram application) and TF32 tensor TFLOP/s (3x that).
"""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
import matfree_b200 as m  # noqa: E402
from matfree_b200 import _device, _lib  # noqa: E402


def main():
    mm = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    ld = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    lib = _lib.load()
    torch.manual_seed(0)
    A = torch.randn(mm, n, device="cuda", dtype=torch.float32) / (mm ** 0.5)
    X = torch.randn(n, ld, device="cuda", dtype=torch.float32)
    op = m.ops.gram(A)
    flops = 4.0 * mm * n * ld
    ref = None
    for variant, tc in ((0, 1), (1, 1), (2, 1), (0, 0)):
        if tc == 0 and flops > 3e12:
            continue  # CUDA-core kernel on the full shape takes too long to be worth timing
        _lib.check(lib.mf_gemm_config(variant, tc))
        W = op.matmat_blocked(X)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            W = op.matmat_blocked(X)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if ref is None:
            # fp64 check on a row sample
            idx = torch.arange(0, n, max(1, n // 64), device="cuda")
            T = A.double() @ X.double()
            ref = A[:, idx].double().T @ T
            del T
        err = float((W[idx].double() - ref).abs().max() / ref.abs().max())
        print(json.dumps({"m": mm, "n": n, "ld": ld, "variant": variant, "tensor_cores": tc,
                          "ms": ms, "fp32_tflops": flops / ms / 1e9,
                          "tf32_tflops": 3 * flops / ms / 1e9 if tc else None,
                          "rel_err_vs_fp64": err}), flush=True)
    lib.mf_gemm_config(0, 1)


if __name__ == "__main__":
    main()

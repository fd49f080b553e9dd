"""Time the x64 dense / Gram probe-block contraction: FP64 tensor cores (DMMA) vs CUDA cores.

usage: python tools/bench_gemm_f64.py [m n ld [reps]]   (defaults: 16384 8192 256)
One JSON line per kernel: ms per Gram matmat, fp64 TFLOP/s (4 m n ld flops per application),
max relative deviation between the two kernels.
"""
import json
import sys

import torch

sys.path.insert(0, ".")
import matfree_b200 as m  # noqa: E402
from matfree_b200 import _lib  # noqa: E402


def main():
    mm = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    ld = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    lib = _lib.load()
    torch.manual_seed(0)
    A = torch.randn(mm, n, device="cuda", dtype=torch.float64) / (mm ** 0.5)
    X = torch.randn(n, ld, device="cuda", dtype=torch.float64)
    op = m.ops.gram(A)
    flops = 4.0 * mm * n * ld
    outs = {}
    for tc in (1, 0):
        _lib.check(lib.mf_gemm_config(0, tc))
        W = op.matmat_blocked(X)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            W = op.matmat_blocked(X)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        outs[tc] = W
        print(json.dumps({"m": mm, "n": n, "ld": ld, "kernel": "dmma" if tc else "cuda_cores",
                          "ms": ms, "fp64_tflops": flops / ms / 1e9}), flush=True)
    ref = A.T @ (A @ X)  # cuBLAS fp64, only as a third opinion
    for tc, W in outs.items():
        print(json.dumps({"kernel": "dmma" if tc else "cuda_cores",
                          "rel_dev_vs_cublas": float((W - ref).abs().max() / ref.abs().max())}), flush=True)
    lib.mf_gemm_config(0, 1)


if __name__ == "__main__":
    main()

"""BASELINE config 4: tridiag_sym with full re-orthogonalisation, depth 100, on the 3-D 7-point
Laplacian 256^3 (n = 16 777 216), rows sharded over the ranks (slabs of grid planes, one plane
of halo per side exchanged over NCCL/NVLink before every product).

  python tools/bench_c4.py                         (1 GPU: the shard is the whole operator)
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_c4.py [--grid 256 --depth 100]

Prints one JSON line (rank 0): ms per decomposition (CUDA events, max over ranks), the
algorithmic bytes of SURVEY.md section 8(d) -- sum_i [4(i+1)+9] n s + matrix -- per GPU and the
fraction of the measured HBM bandwidth, plus a check of the result against the closed-form
spectrum (Ritz values lie inside it; extreme ones converge to its ends).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matfree_b200 as m  # noqa: E402
from matfree_b200 import _lib, _rowshard, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--depth", type=int, default=100)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--planes", type=int, default=0,
                    help="grid planes along the sharded axis (default: --grid); e.g. 32 on one GPU "
                         "reproduces the per-GPU problem of the 8-GPU run")
    ap.add_argument("--no-kernel-timing", action="store_true",
                    help="do not bracket every launch with CUDA events (mf_timing_enable): the ~700 launches "
                         "of a decomposition then run back to back, which is what a user sees")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    g = a.grid
    shape = (a.planes or g, g, g)
    n = shape[0] * g * g
    plane = g * g
    k = a.depth
    r0, r1 = _rowshard.slab_range(n, world, rank, align=plane)
    ip, ix, d = workloads.laplacian_csr_rows(shape, r0, r1, shift=1.0, device=dev)
    nnz_local = int(d.numel())
    op = m.ops.csr_row_sharded(ip, ix, d, n, r0)
    del ix
    torch.cuda.empty_cache()
    # start vector = Rademacher probe 0 of PRNGKey(1): my slab of it, generated in place
    v_full = m.prng.rademacher(m.prng.prng_key(1), shape=(n,), dtype=np.float32)
    v = v_full[r0:r1].clone()
    del v_full
    tri = m.decomp.tridiag_sym(k, reortho="full", materialize=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(a.warmup):
        out = tri(op, v)
    barrier()
    _lib.timing_enable(not a.no_kernel_timing)
    lib = _lib.load()
    l0 = lib.mf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = tri(op, v)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (lib.mf_launch_count() - l0) // a.steps
    per_class = _lib.timing_collect()
    _lib.timing_enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    Q, (diag, off), res, c = out
    route = "peer memory (mf_lanczos_sharded: halo stores + all-reduce fused into the reducing kernels)" \
        if _rowshard._use_peer_memory(None) else "NCCL (Python step loop: send/recv halo, all-reduce per reduction)"
    if getattr(op, "_comm", None) is not None:
        op._comm.check()  # raises if an in-kernel wait timed out
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        s = 4
        nloc = r1 - r0
        matrix = nnz_local * (s + 4) + 4 * (nloc + 1)
        alg = sum((4 * (i + 1) + 9) * nloc * s + matrix for i in range(k))
        T = np.diag(diag.double().cpu().numpy()) + np.diag(off.double().cpu().numpy(), 1) + np.diag(off.double().cpu().numpy(), -1)
        theta = np.linalg.eigvalsh(T)
        lam_1d = 2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1))
        lam_p = 2.0 - 2.0 * np.cos(np.arange(1, shape[0] + 1) * np.pi / (shape[0] + 1))
        lo, hi = 2 * lam_1d.min() + lam_p.min() + 1.0, 2 * lam_1d.max() + lam_p.max() + 1.0
        total_ms = sum(v_[0] for v_ in per_class.values()) or 1.0
        kernels = {c_: {"ms_total_per_decomposition": v_[0] / a.steps, "launches": int(v_[1] // a.steps),
                        "share": v_[0] / total_ms}
                   for c_, v_ in sorted(per_class.items(), key=lambda kv: -kv[1][0])}
        line = {
            "workload": f"C4: tridiag_sym(reortho=full), depth {k}, 3-D 7-pt Laplacian {shape[0]}x{g}x{g} (n={n}) + 1.0*I, fp32, "
                        f"row-sharded x{world} (slabs of {nloc // plane} planes, halo {op.plan.halo_rows} rows)",
            "n_gpus": world, "route": route, "ms_per_decomposition": ms, "steps": a.steps, "warmup": a.warmup,
            "per_launch_event_bracketing": not a.no_kernel_timing,
            "algorithmic_bytes_per_gpu": alg, "achieved_gbs_per_gpu": alg / (ms * 1e-3) / 1e9,
            "hbm_peak_gbs": hbm, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / hbm,
            "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
            "gpu_launches_per_decomposition": int(launches), "kernels": kernels,
            "result": {"ritz_min": float(theta.min()), "ritz_max": float(theta.max()), "spectrum": [lo, hi],
                       "ritz_inside_spectrum": bool(theta.min() >= lo - 1e-4 and theta.max() <= hi + 1e-4),
                       # reortho="full" returns |v| in this slot (matfree/decomp.py:142, an upstream quirk)
                       "init_length_inv": float(c), "expected_init_length_inv": float(np.sqrt(n))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

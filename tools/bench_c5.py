"""BASELINE config 5: funm Lanczos action exp(-t L) v on a synthetic power-law graph Laplacian
(1e7 nodes, ~1e8 non-zeros, CSR), 4096 normal probes, depth 30 (assumed: BASELINE.json does not
state it), two-pass `funm_lanczos_sym` (no stored basis), probe tiles of 256.

  python tools/bench_c5.py [--nodes 10000000 --edges 50000000 --probes 4096 --depth 30]
  (under torchrun: probes are sharded across the ranks, operator replicated)

One JSON line: probe*Lanczos-steps/s (P*k / time; the two passes make 2k matvecs per probe),
per-kernel times, the CSR kernel's achieved bandwidth on this irregular matrix.
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
from matfree_b200 import _lib, _sharding, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=10_000_000)
    ap.add_argument("--edges", type=int, default=50_000_000)
    ap.add_argument("--probes", type=int, default=4096)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--max-tiles", type=int, default=0, help="time only this many tiles (0 = all)")
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
    n, k, ld = a.nodes, a.depth, a.tile
    ip, ix, d, dmax = workloads.powerlaw_laplacian_csr(n, a.edges, device=dev)
    nnz = int(d.numel())
    op = m.ops.csr(ip, ix, d)
    torch.cuda.empty_cache()
    t = 1.0 / dmax
    fun = m.funm.funm_lanczos_sym(m.funm.dense_funm_sym_eigh(("exp", -t)), m.decomp.tridiag_sym(k, reortho="none"))
    p0, p1 = _sharding.shard_range(a.probes, world, rank)
    lib = _lib.load()
    Xb = torch.empty((n, ld), dtype=torch.float32, device=dev)
    key = m.prng.prng_key(1)

    def run_tile(t0):
        npb = min(ld, p1 - t0)
        _lib.check(lib.mf_probe_gen(Xb.data_ptr(), 0, _lib.MF_LAYOUT_BLOCKED, n, ld, t0, npb, int(key[0]),
                                    int(key[1]), _lib.MF_SAMPLER_NORMAL, 0, None,
                                    torch.cuda.current_stream().cuda_stream))
        return fun.blocked(op, Xb, npb), npb

    tiles = list(range(p0, p1, ld))
    if a.max_tiles:
        tiles = tiles[: a.max_tiles]
    out, _ = run_tile(tiles[0])  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.timing_enable(True)
    _lib.timing_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    checksum = 0.0
    for t0 in tiles:
        out, npb = run_tile(t0)
        done += npb
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    per_class = _lib.timing_collect()
    _lib.timing_enable(False)
    # heat-kernel sanity: exp(-tL) preserves the mean of every vector (L 1 = 0) and contracts norms
    col_mean_in = float(Xb[:, 0].double().mean())
    col_mean_out = float(out[:, 0].double().mean())
    norm_ratio = float(out[:, 0].double().norm() / Xb[:, 0].double().norm())
    tot = torch.tensor([float(done), ms], device=dev, dtype=torch.float64)
    if world > 1:
        t_done = tot[:1].clone()
        dist.all_reduce(t_done)
        t_ms = tot[1:].clone()
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        done_all, ms = float(t_done.item()), float(t_ms.item())
    else:
        done_all = float(done)
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        blk = n * ld * 4
        matrix = nnz * 8 + 4 * (n + 1)
        total_ms = sum(v[0] for v in per_class.values()) or 1.0
        kernels = {}
        for c, (kms, cnt) in sorted(per_class.items(), key=lambda kv: -kv[1][0]):
            ent = {"ms_per_launch": kms / cnt, "launches": int(cnt), "share": kms / total_ms}
            if c == "spmm_csr":
                # SURVEY 8(d) counts every X row once (neighbours from L2) -- true for stencils, not for a
                # random graph whose X block (n*ld*4 = 10 GB) is 80x the L2: there every non-zero is a
                # 1 KB gather from DRAM (minus the hub rows that stay cached).  Both are reported.
                ent["achieved_gbs"] = (2 * blk + matrix) / (kms / cnt * 1e-3) / 1e9
                ent["frac_of_hbm_peak"] = ent["achieved_gbs"] / hbm
                gather = nnz * ld * 4 + blk + matrix
                ent["gather_model_gbs"] = gather / (kms / cnt * 1e-3) / 1e9
                ent["gather_model_frac_of_hbm_peak"] = ent["gather_model_gbs"] / hbm
            if c == "lanczos_update":
                ent["achieved_gbs"] = 4 * blk / (kms / cnt * 1e-3) / 1e9
                ent["frac_of_hbm_peak"] = ent["achieved_gbs"] / hbm
            kernels[c] = ent
        # per probe*step of ONE pass: 6 n s + matrix / ld; two passes + the accumulation (3 n s)
        step_bytes = 2 * (6 * n * 4 + matrix / ld) + 3 * n * 4
        value = done_all * k / (ms * 1e-3)
        print(json.dumps({
            "workload": f"C5: exp(-tL)v, power-law graph Laplacian n={n}, nnz={nnz}, d_max={dmax}, t=1/d_max, "
                        f"{int(done_all)} normal probes (of {a.probes}), depth {k} (assumed), two-pass funm_lanczos_sym, tile {ld}",
            "n_gpus": world, "metric": "probe_lanczos_steps_per_sec", "value": value, "ms_total": ms,
            "algorithmic_bytes_per_probe_step": step_bytes,
            "achieved_gbs_per_gpu": value / world * step_bytes / 1e9,
            "frac_of_hbm_peak": value / world * step_bytes / 1e9 / hbm,
            "kernels": kernels,
            "result": {"mean_in": col_mean_in, "mean_out": col_mean_out, "norm_ratio": norm_ratio},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

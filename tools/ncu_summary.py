#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/<name>_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [regex]        > profiles/<name>_full.txt
  python tools/ncu_summary.py wide gpurun_out/prof.ncu-rep [regex]        > profiles/<name>.txt   (all memory-hierarchy counters)
  python tools/ncu_summary.py hotspots gpurun_out/prof.ncu-rep [top]      > profiles/<name>_hotspots.txt
      (needs a capture made with --import-source on: SASS lines by warp-stall samples, with the stall reasons)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.split("::")[-1].strip()[:70]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary of {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} ms total (cold-cache, serialised)")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'share':>6s} {'avg ms':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 2e-4:
            continue
        print(f"{k:70s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot:6.3f} {v[1] / v[0]:9.4f}")


def wide(path, pattern=None):
    """Every metric of the raw page whose name matches WIDE (memory hierarchy, L1 data pipe,
    occupancy, stalls): used on the GPU box, where only the text summary is brought back."""
    keep = re.compile(r"^(gpu__time|dram__|lts__t_bytes|lts__throughput|lts__t_sector|lts__t_sectors_srcunit_tex"
                      r"|l1tex__|sm__throughput|sm__warps_active|sm__inst_executed_pipe_(lsu|fp64|alu|fma)"
                      r"|smsp__inst_executed\.sum|launch__(registers|grid|block|occupancy|shared)|smsp__average_warp"
                      r"|smsp__cycles_active\.avg|sm__cycles_elapsed\.max|smsp__warp_issue_stalled.*_per_warp_active)")
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full, all memory-hierarchy metrics of {path}")
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        print(f"\n== {name}  (launch id {r[idx['ID']]})")
        for h in hdr:
            if keep.search(h):
                print(f"  {h:95s} {r[idx[h]]:>18s} {units[idx[h]]}")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of {path}")
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        print(f"\n== {name}  (launch id {r[idx['ID']]})")
        for m in KEEP:
            if m in idx:
                print(f"  {m:85s} {r[idx[m]]:>18s} {units[idx[m]]}")


def hotspots(path, top=40):
    """SASS lines of a `--set full --import-source on` capture ordered by warp-stall samples."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h0 = next(i for i, r in enumerate(rows) if any("Sampl" in c for c in r))
    hdr = rows[h0]
    key = next(i for i, c in enumerate(hdr) if "Sampl" in c)
    src = hdr.index("Source") if "Source" in hdr else 1

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0

    body = [r for r in rows[h0 + 1:] if len(r) == len(hdr)]
    tot = sum(num(r[key]) for r in body) or 1.0
    print(f"# warp-stall sampling hot spots of {path}: % of samples, SASS, stall reasons (samples)")
    for r in sorted(body, key=lambda r: -num(r[key]))[:top]:
        why = " ".join(f"{h}={r[i]}" for i, h in enumerate(hdr)
                       if h.startswith("stall") and "Not" not in h and num(r[i]) > 0)
        print(f"{100 * num(r[key]) / tot:5.1f} {r[src][:90]} | {why[:200]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "hotspots":
        hotspots(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
    elif sys.argv[1] == "wide":
        wide(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

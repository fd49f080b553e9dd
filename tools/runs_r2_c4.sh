#!/bin/bash
# C4 after the deep-ring fused CGS kernel and the 4-CTA bound for the narrow 7-point SpMM:
# parity, then 1-GPU full size and the 32-plane slab (the per-GPU problem of the 8-GPU run)
timeout 900 python -m pytest tests -m gpu -q -k "full or c4 or tridiag or hessenberg or reorth or cgs or eig or arnoldi or single_vector" 2>&1 | tail -3
out=gpurun_out/r2zi_c4.jsonl
: > $out
timeout 300 python tools/bench_c4.py --steps 3 --warmup 1 --no-kernel-timing >> $out 2>>gpurun_out/r2zi.err
timeout 300 python tools/bench_c4.py --steps 2 --warmup 1 >> $out 2>>gpurun_out/r2zi.err
timeout 300 python tools/bench_c4.py --steps 3 --warmup 1 --planes 32 --no-kernel-timing >> $out 2>>gpurun_out/r2zi.err
timeout 300 python tools/bench_c4.py --steps 2 --warmup 1 --planes 32 >> $out 2>>gpurun_out/r2zi.err
MF_SPMM_ROW_THREAD=0 timeout 300 python tools/bench_c4.py --steps 2 --warmup 1 >> $out 2>>gpurun_out/r2zi.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2zi_c4.jsonl"):
    d=json.loads(ln)
    print(d["workload"][50:80], d["ms_per_decomposition"], round(d["frac_of_hbm_peak"],3), d["per_launch_event_bracketing"], d["result"]["ritz_inside_spectrum"], d["result"]["ritz_min"], d["result"]["ritz_max"])
    for k,v in d["kernels"].items(): print("    ",k, round(v["ms_total_per_decomposition"],2), v["launches"])
PY
tail -3 gpurun_out/r2zi.err

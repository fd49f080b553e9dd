#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_cgs_fused.py tests/test_gpu_rowshard.py -x -q ) > gpurun_out/r2a_tests.log 2>&1
tail -5 gpurun_out/r2a_tests.log
( timeout 300 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r2a_c4_1gpu.json 2> gpurun_out/r2a_c4.err
python - <<'PY'
import json
for l in open("gpurun_out/r2a_c4_1gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_decomposition"],2), round(d["frac_of_hbm_peak"],3), d["gpu_launches_per_decomposition"], {k:round(v["ms_total_per_decomposition"],2) for k,v in d["kernels"].items()}, d["result"]["ritz_min"], d["result"]["ritz_max"])
PY
tail -2 gpurun_out/r2a_c4.err

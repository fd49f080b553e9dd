#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cgs_fused.py tests/test_gpu_rowshard.py tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q ) > gpurun_out/r1q_tests.log 2>&1
tail -25 gpurun_out/r1q_tests.log
for env in "" "MF_CGS_FUSED_OFF=1"; do
( env $env timeout 600 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r1q_c4_1gpu_${env:-fused}.json 2> gpurun_out/r1q_c4.err
( env $env timeout 600 python tools/bench_c4.py --planes 32 --steps 3 --warmup 1 ) > gpurun_out/r1q_c4_p32_${env:-fused}.json 2>> gpurun_out/r1q_c4.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r1q_c4_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_decomposition"],2), round(d["frac_of_hbm_peak"],3), d["gpu_launches_per_decomposition"], {k:round(v["ms_total_per_decomposition"],2) for k,v in d["kernels"].items()}, d["result"]["ritz_min"], d["result"]["ritz_max"])
PY
tail -3 gpurun_out/r1q_c4.err

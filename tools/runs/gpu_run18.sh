#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowshard.py tests/test_gpu_configs.py tests/test_gpu_gemm_dmma.py -x -q ) > gpurun_out/r1i_tests.log 2>&1
tail -3 gpurun_out/r1i_tests.log
( timeout 600 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r1i_c4_1gpu.json 2> gpurun_out/r1i_c4_1gpu.err
python - <<'PY'
import json
for f in ["gpurun_out/r1i_c4_1gpu.json"]:
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_decomposition"],2), round(d["frac_of_hbm_peak"],3), {k:round(v["ms_total_per_decomposition"],2) for k,v in d["kernels"].items()})
PY
( timeout 600 python tools/bench_c4.py --planes 32 --steps 3 --warmup 1 ) > gpurun_out/r1i_c4_p32.json 2> gpurun_out/r1i_c4_p32.err
python - <<'PY'
import json
for f in ["gpurun_out/r1i_c4_p32.json"]:
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_decomposition"],2), round(d["frac_of_hbm_peak"],3), {k:round(v["ms_total_per_decomposition"],2) for k,v in d["kernels"].items()})
PY

#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cgs_fused.py -x -q ) > gpurun_out/r1q_tests.log 2>&1
tail -5 gpurun_out/r1q_tests.log
for env in "MF_CGS_FUSED=1" "MF_X=0"; do
( env $env timeout 600 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r1v_c4_1gpu_${env}.json 2> gpurun_out/r1v_c4.err
( env $env timeout 600 python tools/bench_c4.py --planes 32 --steps 3 --warmup 1 ) > gpurun_out/r1v_c4_p32_${env}.json 2>> gpurun_out/r1v_c4.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r1v_c4_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_decomposition"],2), round(d["frac_of_hbm_peak"],3), d["gpu_launches_per_decomposition"], {k:round(v["ms_total_per_decomposition"],2) for k,v in d["kernels"].items()})
PY
tail -3 gpurun_out/r1v_c4.err

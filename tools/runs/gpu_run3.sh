#!/bin/bash
# SpMM tuning: L2 prefetch on/off x rows per chunk; ncu full capture of the SpMM.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['kernels']
        print('value', round(d['value']), {n: round(v['ms_per_launch'], 3) for n, v in k.items()})
"
}
{
run MF_SPMM_PREFETCH=0 MF_SPMM_ROWS=64
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=16
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=32
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=64
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=128
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=256
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=64 MF_SPMM_GROUP=4
run MF_SPMM_PREFETCH=1 MF_SPMM_ROWS=32 MF_SPMM_GROUP=4
} > gpurun_out/tune_spmm.log 2>&1
cat gpurun_out/tune_spmm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr' -s 8 -c 2 \
  -o gpurun_out/prof_r1b -f python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log

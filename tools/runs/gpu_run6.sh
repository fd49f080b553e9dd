#!/bin/bash
# SpMM v4 (cp.async ring, static chunks + throttle): correctness then tuning.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline"
run() {
  echo "== $*"
  env "$@" timeout 300 $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['kernels']
        print('value', round(d['value']), {n: round(v['ms_per_launch'], 3) for n, v in k.items()})
"
}
prof() {
  env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:spmm_csr -s 6 -c 1 --csv $B 2>/dev/null | grep '^"' | python -c "
import sys, csv
for r in csv.reader(sys.stdin):
    if r[0] != 'ID': print('   ncu', r[-3], r[-2], r[-1])
"
}
{
run MF_SPMM_ROWS=64
prof MF_SPMM_ROWS=64
run MF_SPMM_ROWS=32
run MF_SPMM_ROWS=128
run MF_SPMM_ROWS=64 MF_SPMM_NSEG=3
run MF_SPMM_ROWS=64 MF_SPMM_NSEG=2
run MF_SPMM_ROWS=64 MF_SPMM_THROTTLE=0
prof MF_SPMM_ROWS=64 MF_SPMM_THROTTLE=0
run MF_SPMM_ROWS=64 MF_SPMM_SLACK=16
run MF_SPMM_ROWS=64 MF_SPMM_SLACK=592
run MF_SPMM_ROWS=64 MF_SPMM_PREFETCH=1
} > gpurun_out/tune_spmm4.log 2>&1
cat gpurun_out/tune_spmm4.log

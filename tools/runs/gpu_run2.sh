#!/bin/bash
# Second GPU pass: parity tests on the fused-finalize / staged SpMM build, sanitizer spot checks, SpMM tuning sweep, bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "slq_per_probe and float32" ) > gpurun_out/racecheck.log 2>&1
tail -4 gpurun_out/racecheck.log
( timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "slq_per_probe or matmat_csr or tridiag_decomposition" ) > gpurun_out/memcheck.log 2>&1
tail -4 gpurun_out/memcheck.log
for rows in 32 64 128 256; do for grp in 4 8; do
  echo "== MF_SPMM_ROWS=$rows MF_SPMM_GROUP=$grp"
  MF_SPMM_ROWS=$rows MF_SPMM_GROUP=$grp timeout 300 python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['kernels']
        print('value', round(d['value']), {n: round(v['ms_per_launch'], 3) for n, v in k.items()})
"
done; done > gpurun_out/tune_spmm.log 2>&1
cat gpurun_out/tune_spmm.log
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r1n_variants.txt
for lib in libmatfree_b200_prev.so libmatfree_b200.so libmatfree_b200_ctas3.so; do
  export MF_LIB_PATH=$PWD/matfree_b200/_lib/$lib
  echo "== $lib" >> gpurun_out/r1n_variants.txt
  timeout 120 python tools/bench_spmm.py 4096 256 10 >> gpurun_out/r1n_variants.txt 2>&1
  timeout 300 python bench.py --steps 1 --warmup 3 --probes-per-gpu 256 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'value': d['value'], 'kernels': {k:(round(v['ms_per_launch'],3)) for k,v in d['kernels'].items()}}))" >> gpurun_out/r1n_variants.txt 2>&1
done
cat gpurun_out/r1n_variants.txt

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r1s_variants.txt
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q ) 2>&1 | tail -3 >> gpurun_out/r1s_variants.txt
for lib in libmatfree_b200_prev.so libmatfree_b200.so libmatfree_b200_prev.so libmatfree_b200.so; do
  export MF_LIB_PATH=$PWD/matfree_b200/_lib/$lib
  echo "== $lib" >> gpurun_out/r1s_variants.txt
  timeout 300 python bench.py --steps 2 --warmup 3 --probes-per-gpu 256 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'value': d['value'], 'sm_mhz': d['clocks']['sm_mhz'], 'kernels': {k:(round(v['ms_per_launch'],3)) for k,v in d['kernels'].items()}}))" >> gpurun_out/r1s_variants.txt 2>&1
done
cat gpurun_out/r1s_variants.txt

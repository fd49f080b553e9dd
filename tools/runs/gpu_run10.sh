#!/bin/bash
# row-sharded path on one GPU + C1 tests + C4 at N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowshard.py tests/test_gpu_configs.py tests/test_gpu_parity.py -q 2>&1 | tail -25 > gpurun_out/r10_tests.log
cat gpurun_out/r10_tests.log
( time timeout 900 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r10_c4_1gpu.json 2> gpurun_out/r10_c4_1gpu.err
tail -c 2500 gpurun_out/r10_c4_1gpu.json; tail -5 gpurun_out/r10_c4_1gpu.err

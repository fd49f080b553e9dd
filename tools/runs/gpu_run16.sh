#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowshard.py tests/test_gpu_configs.py tests/test_gpu_gemm_dmma.py -x -q ) > gpurun_out/r1g_tests.log 2>&1
tail -3 gpurun_out/r1g_tests.log
( timeout 600 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r1g_c4_1gpu.json 2> gpurun_out/r1g_c4_1gpu.err
tail -c 1400 gpurun_out/r1g_c4_1gpu.json | head -c 900; tail -3 gpurun_out/r1g_c4_1gpu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'reorth_dots|reorth_update' -s 150 -c 2 \
  -o gpurun_out/r1g_c4_prof -f python tools/bench_c4.py --depth 40 --steps 1 --warmup 1 > gpurun_out/r1g_ncu.log 2>&1
tail -2 gpurun_out/r1g_ncu.log

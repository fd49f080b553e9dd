#!/bin/bash
# regression + ncu of the CGS kernels of C4 (1 GPU, depth 40: launches at steps 30..39)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1f_tests.log 2>&1
tail -3 gpurun_out/r1f_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'reorth_dots|reorth_update' -s 150 -c 4 \
  -o gpurun_out/r1f_c4_prof -f python tools/bench_c4.py --depth 40 --steps 1 --warmup 1 > gpurun_out/r1f_ncu.log 2>&1
tail -2 gpurun_out/r1f_ncu.log

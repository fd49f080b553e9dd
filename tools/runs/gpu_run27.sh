#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cgs_update_dots' -s 60 -c 1 \
  -o gpurun_out/r1z_cgs_prof -f python tools/bench_c4.py --depth 80 --steps 1 --warmup 0 > gpurun_out/r1z_ncu.log 2>&1
tail -2 gpurun_out/r1z_ncu.log

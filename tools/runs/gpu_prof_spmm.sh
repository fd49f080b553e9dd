#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr' -s 8 -c 1 \
  -o gpurun_out/prof_spmm -f python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log

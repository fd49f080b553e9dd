#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr' -s 4 -c 1 \
  -o gpurun_out/r1m_spmm_prof -f python tools/bench_spmm.py 4096 256 3 > gpurun_out/r1m_ncu.log 2>&1
tail -2 gpurun_out/r1m_ncu.log

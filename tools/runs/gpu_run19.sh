#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r1j_spmm_sweep.jsonl
for cfg in "" "MF_SPMM_PFD=1" "MF_SPMM_PFD=2" "MF_SPMM_PFD=3" "MF_SPMM_PFD=4" "MF_SPMM_PFD=6" "MF_SPMM_PFD=8" "MF_SPMM_PFD=12" "MF_SPMM_PFD=16" \
           "MF_SPMM_ROWS=32" "MF_SPMM_ROWS=32 MF_SPMM_PFD=4" "MF_SPMM_ROWS=128 MF_SPMM_PFD=4" "MF_SPMM_ROWS=32 MF_SPMM_PFD=8"; do
  env $cfg timeout 120 python tools/bench_spmm.py >> gpurun_out/r1j_spmm_sweep.jsonl 2>> gpurun_out/r1j_spmm_sweep.err
done
cat gpurun_out/r1j_spmm_sweep.jsonl

#!/bin/bash
# ncu --set full of the CSR product inside the Lanczos step: 2-D (gather and band kernels) and 3-D (tile 256 and 64)
B="--steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 256"
ncu --set full --clock-control none --import-source on -k regex:spmm -s 6 -c 2 -o gpurun_out/r2f_spmm_2d_gather -f python bench.py $B > /dev/null 2> gpurun_out/r2f_ncu.err
MF_SPMM_STRIP=1 ncu --set full --clock-control none --import-source on -k regex:spmm -s 6 -c 2 -o gpurun_out/r2f_spmm_2d_band -f python bench.py $B > /dev/null 2>> gpurun_out/r2f_ncu.err
ncu --set full --clock-control none --import-source on -k regex:spmm -s 6 -c 2 -o gpurun_out/r2f_spmm_3d_t256 -f python bench.py --workload c2-3d $B > /dev/null 2>> gpurun_out/r2f_ncu.err
ncu --set full --clock-control none --import-source on -k regex:spmm -s 6 -c 2 -o gpurun_out/r2f_spmm_3d_t64 -f python bench.py --workload c2-3d --tile 64 $B > /dev/null 2>> gpurun_out/r2f_ncu.err
ls -la gpurun_out/*.ncu-rep | tail -5

#!/bin/bash
# ncu --set full of the CSR product inside the Lanczos step: 2-D (gather and band kernels) and 3-D
# (tile 256 and 64).  The reports stay on the box (too large); only text summaries come back.
B="--steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 256"
T=/tmp/ncu_r2
mkdir -p $T
run() {  # name, env, extra bench args
  env $2 ncu --set full --clock-control none -k regex:spmm -s 6 -c 1 -o $T/$1 -f python bench.py $3 $B > /dev/null 2>> gpurun_out/r2f_ncu.err
  python tools/ncu_summary.py wide $T/$1.ncu-rep > gpurun_out/r2f_$1.txt 2>> gpurun_out/r2f_ncu.err
}
run spmm_2d_gather "MF_X=0" ""
run spmm_2d_band "MF_SPMM_STRIP=1" ""
run spmm_3d_t256 "MF_X=0" "--workload c2-3d"
run spmm_3d_t64 "MF_X=0" "--workload c2-3d --tile 64"
wc -l gpurun_out/r2f_*.txt

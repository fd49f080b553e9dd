#!/bin/bash
# ncu --set full of the CSR product on the 3-D target inside the Lanczos step: hybrid TMA band kernel and row-group kernel
B="--workload c2-3d --steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 256"
T=/tmp/ncu_r2; mkdir -p $T
run() {  # name, env
  env $2 timeout 600 ncu --set full --clock-control none -k regex:spmm -s 6 -c 1 -o $T/$1 -f python bench.py $B > /dev/null 2>> gpurun_out/r2v.err
  python tools/ncu_summary.py wide $T/$1.ncu-rep > gpurun_out/r2v_$1.txt 2>> gpurun_out/r2v.err
  echo == $1
  grep -E "gpu__time_duration.sum|l1tex__data_pipe_lsu_wavefronts.avg.pct|smsp__inst_executed.sum |dram__bytes_read.sum |dram__bytes_write.sum |lts__t_sector_hit_rate.pct|smsp__average_warps_issue_stalled_(barrier|long|short|wait|mio|sleeping|selected|not_sel|lg).*ratio|sm__warps_active.avg.pct|launch__grid_size|launch__registers_per_thread " gpurun_out/r2v_$1.txt | cut -c1-150
}
run spmm_3d_tma "MF_SPMM_TMA=2"
run spmm_3d_gather "MF_SPMM_TMA=0"
tail -3 gpurun_out/r2v.err

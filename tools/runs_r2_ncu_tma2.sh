#!/bin/bash
# ncu --set full of the barrier-free TMA band kernel (C2, in the Lanczos step) + warp-stall hot spots by SASS line
B="--steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 256"
T=/tmp/ncu_r2; mkdir -p $T
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmm -s 6 -c 1 -o $T/spmm_2d_tma2 -f python bench.py $B > /dev/null 2>> gpurun_out/r2zf.err
python tools/ncu_summary.py wide $T/spmm_2d_tma2.ncu-rep > gpurun_out/r2zf_spmm_2d_tma.txt 2>> gpurun_out/r2zf.err
python tools/ncu_summary.py hotspots $T/spmm_2d_tma2.ncu-rep > gpurun_out/r2zf_spmm_2d_tma_hotspots.txt 2>> gpurun_out/ncu_hotspots.err
grep -E "gpu__time_duration.sum|l1tex__data_pipe_lsu_wavefronts.avg.pct|smsp__inst_executed.sum |dram__bytes_read.sum |dram__bytes_write.sum |lts__throughput.avg|smsp__average_warps_issue_stalled_(barrier|long|short|wait|mio|sleeping|selected|not_sel|branch|no_inst).*ratio|sm__warps_active.avg.pct|l1tex__m_xbar2l1tex_read_bytes.sum " gpurun_out/r2zf_spmm_2d_tma.txt | cut -c1-150
head -30 gpurun_out/r2zf_spmm_2d_tma_hotspots.txt | cut -c1-300

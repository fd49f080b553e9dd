#!/bin/bash
# ncu --set full of the default CSR product on the 3-D target inside the Lanczos step (chunked 7-diagonal TMA kernel)
B="--workload c2-3d --steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 256"
T=/tmp/ncu_r2; mkdir -p $T
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmm -s 6 -c 1 -o $T/spmm_3d -f python bench.py $B > /dev/null 2>> gpurun_out/r2zm.err
python tools/ncu_summary.py wide $T/spmm_3d.ncu-rep > gpurun_out/r2zm_spmm_3d_tma.txt 2>> gpurun_out/r2zm.err
python tools/ncu_summary.py hotspots $T/spmm_3d.ncu-rep > gpurun_out/r2zm_spmm_3d_tma_hotspots.txt 2>> gpurun_out/ncu_hotspots.err
grep -E "gpu__time_duration.sum|l1tex__data_pipe_lsu_wavefronts.avg.pct|smsp__inst_executed.sum |dram__bytes_read.sum |dram__bytes_write.sum |lts__throughput.avg|lts__t_sector_hit_rate.pct|smsp__average_warps_issue_stalled_(barrier|long|short|wait|mio|sleeping|selected|not_sel|branch|no_inst|lg).*ratio|sm__warps_active.avg.pct|l1tex__m_xbar2l1tex_read_bytes.sum " gpurun_out/r2zm_spmm_3d_tma.txt | cut -c1-150
head -24 gpurun_out/r2zm_spmm_3d_tma_hotspots.txt | cut -c1-260

#!/bin/bash
# ncu --set full of the fused CGS kernel (C4, 1 GPU, launch 80: nq = 81) and of the plain update
# kernel beside it; text summaries only (the reports stay on the box)
T=/tmp/ncu_r2; mkdir -p $T
A="tools/bench_c4.py --steps 1 --warmup 0 --no-kernel-timing"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cgs_update -s 80 -c 1 -o $T/cgs -f python $A > /dev/null 2>> gpurun_out/r2t.err
python tools/ncu_summary.py wide $T/cgs.ncu-rep > gpurun_out/r2t_cgs.txt 2>> gpurun_out/r2t.err
python tools/ncu_summary.py hotspots $T/cgs.ncu-rep > gpurun_out/r2t_cgs_source.txt 2>> gpurun_out/ncu_hotspots.err
#false && timeout 600 ncu --set full --clock-control none -k regex:reorth_update_kernel -s 160 -c 1 -o $T/upd -f python $A > /dev/null 2>> gpurun_out/r2t.err
#false && python tools/ncu_summary.py wide $T/upd.ncu-rep > gpurun_out/r2t_upd.txt 2>> gpurun_out/r2t.err
for f in gpurun_out/r2t_cgs.txt gpurun_out/r2t_upd.txt; do echo == $f; grep -E "gpu__time_duration.sum|l1tex__data_pipe_lsu_wavefronts.avg.pct|l1tex__data_bank_conflicts|smsp__inst_executed.sum |dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct|smsp__average_warps_issue_stalled_.*ratio|sm__warps_active.avg.pct|lts__t_sectors_srcunit_tex_op_read.sum |l1tex__throughput.avg.pct|launch__shared_mem_per_block_dynamic|launch__grid_size" $f | cut -c1-150; done
head -70 gpurun_out/r2t_cgs_source.txt | cut -c1-380
tail -3 gpurun_out/r2t.err

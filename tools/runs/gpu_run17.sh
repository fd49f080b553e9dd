#!/bin/bash
# per-launch durations of the C4 kernels at the per-GPU size of the 8-GPU run (32 planes of 256^2 on one GPU)
mkdir -p gpurun_out
( timeout 600 python tools/bench_c4.py --planes 32 --steps 3 --warmup 1 ) > gpurun_out/r1h_c4_p32.json 2> gpurun_out/r1h_c4_p32.err
tail -c 1200 gpurun_out/r1h_c4_p32.json | head -c 800; tail -3 gpurun_out/r1h_c4_p32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/r1h_c4_p32_launches.csv python tools/bench_c4.py --planes 32 --steps 1 --warmup 0 > gpurun_out/r1h_ncu_list.log 2>&1
tail -2 gpurun_out/r1h_ncu_list.log

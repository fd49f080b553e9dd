#!/bin/bash
# single-GPU regression + C4 on one GPU (new narrow-tile dots kernel)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1f_tests.log 2>&1
tail -5 gpurun_out/r1f_tests.log
( timeout 600 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r1f_c4_1gpu.json 2> gpurun_out/r1f_c4_1gpu.err
tail -c 1800 gpurun_out/r1f_c4_1gpu.json; tail -3 gpurun_out/r1f_c4_1gpu.err

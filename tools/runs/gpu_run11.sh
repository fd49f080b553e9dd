#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r11_tests.log
cat gpurun_out/r11_tests.log
( time timeout 900 python tools/bench_c5.py --max-tiles 2 ) > gpurun_out/r11_c5.json 2> gpurun_out/r11_c5.err
tail -c 2500 gpurun_out/r11_c5.json; tail -5 gpurun_out/r11_c5.err

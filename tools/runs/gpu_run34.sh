#!/bin/bash
# final check of the default bench command (r2b)
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 1500 gpurun_out/r2b_bench.json; tail -4 gpurun_out/r2b_bench.err

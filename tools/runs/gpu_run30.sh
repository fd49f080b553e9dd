#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -k "c4_full" ) > gpurun_out/r1u_tests.log 2>&1
tail -30 gpurun_out/r1u_tests.log

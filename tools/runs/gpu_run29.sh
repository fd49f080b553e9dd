#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_eig.py -x -q ) > gpurun_out/r1t_tests.log 2>&1
tail -30 gpurun_out/r1t_tests.log

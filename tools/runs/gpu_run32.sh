#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_golden.py -x -q ) > gpurun_out/r1y_tests.log 2>&1
tail -25 gpurun_out/r1y_tests.log

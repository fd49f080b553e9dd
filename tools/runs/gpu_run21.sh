#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_bidiag.py -x -q ) > gpurun_out/r1l_tests.log 2>&1
tail -40 gpurun_out/r1l_tests.log

#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_hutchinson.py -x -q ) > gpurun_out/r1k_tests.log 2>&1
tail -25 gpurun_out/r1k_tests.log

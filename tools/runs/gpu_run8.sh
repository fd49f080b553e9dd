#!/bin/bash
# tcgen05 GEMM bring-up: parity tests first (bounded), then timing on C3-like shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q 2>&1 | tail -30 > gpurun_out/r8_gemm_tests.log
cat gpurun_out/r8_gemm_tests.log
timeout 300 python tools/bench_gemm.py 8192 4096 256 3 2>&1 | tail -8 > gpurun_out/r8_gemm_small.log
cat gpurun_out/r8_gemm_small.log
timeout 600 python tools/bench_gemm.py 65536 16384 256 3 2>&1 | tail -8 > gpurun_out/r8_gemm_c3.log
cat gpurun_out/r8_gemm_c3.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/r8_parity.log
cat gpurun_out/r8_parity.log

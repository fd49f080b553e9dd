#!/bin/bash
# r1e single-GPU pass: DMMA tests + micro-bench, row-shard tests through the native sharded driver, C4 on 1 GPU
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gemm_dmma.py tests/test_gpu_rowshard.py tests/test_gpu_parity.py -x -q ) > gpurun_out/r1e_tests.log 2>&1
tail -15 gpurun_out/r1e_tests.log
( timeout 300 python tools/bench_gemm_f64.py ) > gpurun_out/r1e_gemm_f64.jsonl 2>&1
cat gpurun_out/r1e_gemm_f64.jsonl | tail -6
( timeout 600 python tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/r1e_c4_1gpu.json 2> gpurun_out/r1e_c4_1gpu.err
tail -c 2000 gpurun_out/r1e_c4_1gpu.json; tail -3 gpurun_out/r1e_c4_1gpu.err

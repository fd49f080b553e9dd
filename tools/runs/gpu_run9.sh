#!/bin/bash
# tcgen05 path: config-level parity, C3 bench line (supplementary), ncu of the GEMM kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_gemm_tc.py -q 2>&1 | tail -25 > gpurun_out/r9_tests.log
cat gpurun_out/r9_tests.log
( time timeout 1200 python bench.py --workload c3 --steps 2 --warmup 3 ) > gpurun_out/r9_bench_c3.json 2> gpurun_out/r9_bench_c3.err
tail -c 3000 gpurun_out/r9_bench_c3.json; tail -5 gpurun_out/r9_bench_c3.err
( timeout 600 python bench.py --workload c3 --impl reference --steps 1 --warmup 1 ) > gpurun_out/r9_bench_c3_ref.json 2> gpurun_out/r9_bench_c3_ref.err
tail -c 600 gpurun_out/r9_bench_c3_ref.json; tail -3 gpurun_out/r9_bench_c3_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3' -s 2 -c 2 \
  -o gpurun_out/r9_gemm_prof -f python tools/bench_gemm.py 65536 16384 256 1 > gpurun_out/r9_ncu_gemm.log 2>&1
tail -3 gpurun_out/r9_ncu_gemm.log

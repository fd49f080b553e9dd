#!/bin/bash
# compute-sanitizer over the kernels added this session (memcheck + racecheck on small cases)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q \
    tests/test_gpu_rowshard.py tests/test_gpu_hutchinson.py "tests/test_gpu_gemm_dmma.py::test_gram_dmma_matches_numpy" \
    "tests/test_gpu_bidiag.py::test_bidiag_decomposition_is_satisfied" ) > gpurun_out/r1p_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r1p_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -x -q \
    tests/test_gpu_rowshard.py "tests/test_gpu_gemm_dmma.py::test_gram_dmma_matches_numpy" ) > gpurun_out/r1p_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/r1p_racecheck.log

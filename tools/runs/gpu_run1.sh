#!/bin/bash
# First GPU pass: parity tests, smoke, bench, ncu launch list + full captures of the two HBM kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr|lanczos_update' -s 8 -c 4 \
  -o gpurun_out/prof_r1 -f python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

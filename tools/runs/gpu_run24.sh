#!/bin/bash
# r1o validation pass: all GPU tests, smoke, bench (both arms), ncu launch list + full capture of the three HBM kernels
TAG=r1o
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -1 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
( timeout 600 python bench.py --impl reference ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 300 gpurun_out/${TAG}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr|lanczos_update|probe_gen' -s 0 -c 5 \
  -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --profile --probes-per-gpu 256 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log

#!/bin/bash
# r1d validation pass: GPU parity tests, smoke, bench (both arms), C5 bench, ncu launch list of the bench command.
TAG=r1d
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -1 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
( timeout 600 python bench.py --impl reference ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 400 gpurun_out/${TAG}_bench_ref.json
( time timeout 900 python tools/bench_c5.py --max-tiles 2 ) > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err
tail -c 2500 gpurun_out/${TAG}_c5.json; tail -5 gpurun_out/${TAG}_c5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_list.log

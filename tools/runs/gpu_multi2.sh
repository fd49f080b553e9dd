#!/bin/bash
# multi-GPU pass: NCCL parity tests (probe sharding + row sharding), C4 row-sharded, C2 probe-sharded
# usage: tools/runs/gpu_multi2.sh <ngpus>
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15 > gpurun_out/m${N}_tests.log
cat gpurun_out/m${N}_tests.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
   tools/bench_c4.py --steps 2 --warmup 1 ) > gpurun_out/m${N}_c4.json 2> gpurun_out/m${N}_c4.err
tail -c 2500 gpurun_out/m${N}_c4.json; tail -3 gpurun_out/m${N}_c4.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 \
   bench.py --gpus $N --steps 2 --warmup 3 --no-e2e ) > gpurun_out/m${N}_bench_c2.json 2> gpurun_out/m${N}_bench_c2.err
tail -c 1200 gpurun_out/m${N}_bench_c2.json; tail -3 gpurun_out/m${N}_bench_c2.err
if [ "${MF_C3:-0}" = "1" ]; then
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 \
   bench.py --workload c3 --gpus $N --steps 2 --warmup 3 --no-e2e ) > gpurun_out/m${N}_bench_c3.json 2> gpurun_out/m${N}_bench_c3.err
tail -c 900 gpurun_out/m${N}_bench_c3.json; tail -3 gpurun_out/m${N}_bench_c3.err
fi

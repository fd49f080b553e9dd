#!/bin/bash
# multi-GPU pass for the peer-memory path: NCCL/peer parity tests, C4 row-sharded (peer route and NCCL route)
# usage: tools/runs/gpu_multi3.sh <ngpus> [tag]
N=${1:-2}
TAG=${2:-r1e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_m${N}_tests.log
cat gpurun_out/${TAG}_m${N}_tests.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
   tools/bench_c4.py --steps 3 --warmup 1 ) > gpurun_out/${TAG}_m${N}_c4.json 2> gpurun_out/${TAG}_m${N}_c4.err
tail -c 2500 gpurun_out/${TAG}_m${N}_c4.json; tail -3 gpurun_out/${TAG}_m${N}_c4.err
( MF_ROWSHARD_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 \
   tools/bench_c4.py --steps 3 --warmup 1 ) > gpurun_out/${TAG}_m${N}_c4_nccl.json 2> gpurun_out/${TAG}_m${N}_c4_nccl.err
tail -c 600 gpurun_out/${TAG}_m${N}_c4_nccl.json; tail -3 gpurun_out/${TAG}_m${N}_c4_nccl.err

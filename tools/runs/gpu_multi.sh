#!/bin/bash
# Multi-GPU pass: NCCL parity test + bench at N = $1 GPUs.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/multi_pytest.log 2>&1
tail -5 gpurun_out/multi_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/multi_bench_$N.json 2> gpurun_out/multi_bench_$N.err
tail -c 1500 gpurun_out/multi_bench_$N.json; tail -5 gpurun_out/multi_bench_$N.err

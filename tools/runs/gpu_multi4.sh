#!/bin/bash
# multi-GPU pass: parity tests (probe sharding, row sharding via peer memory and NCCL), C4 peer route (fused CGS on/off),
# C4 NCCL route, C2 probe-sharded bench
# usage: tools/runs/gpu_multi4.sh <ngpus> [tag]
N=${1:-8}
TAG=${2:-r1x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_m${N}_tests.log
cat gpurun_out/${TAG}_m${N}_tests.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
   tools/bench_c4.py --steps 3 --warmup 1 ) > gpurun_out/${TAG}_m${N}_c4.json 2> gpurun_out/${TAG}_m${N}_c4.err
( MF_CGS_FUSED_OFF=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 \
   tools/bench_c4.py --steps 3 --warmup 1 ) > gpurun_out/${TAG}_m${N}_c4_unfused.json 2> gpurun_out/${TAG}_m${N}_c4_unfused.err
( MF_ROWSHARD_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 \
   tools/bench_c4.py --steps 3 --warmup 1 ) > gpurun_out/${TAG}_m${N}_c4_nccl.json 2> gpurun_out/${TAG}_m${N}_c4_nccl.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29574 \
   bench.py --gpus $N --steps 2 --warmup 3 ) > gpurun_out/${TAG}_m${N}_bench_c2.json 2> gpurun_out/${TAG}_m${N}_bench_c2.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_m${N}_c4*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_decomposition"],2), round(d["frac_of_hbm_peak"],3), d["gpu_launches_per_decomposition"], {k:round(v["ms_total_per_decomposition"],2) for k,v in d["kernels"].items()})
for l in open("gpurun_out/${TAG}_m${N}_bench_c2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("C2", d["n_gpus"], d["value"], d["ms_per_step"], d.get("e2e",{}) and d["e2e"].get("value"), d["clocks"])
PY
tail -2 gpurun_out/${TAG}_m${N}_c4.err gpurun_out/${TAG}_m${N}_bench_c2.err

#!/bin/bash
# band kernel with L1 prefetch (2-D), band kernel on the 3-D workload (blocked row order)
python tools/bench_spmm.py --configs 0:64:2:3,1:64:0:3,1:64:-1:3,1:64:-2:3,1:64:-3:3,1:64:-4:3,1:32:-2:3,1:128:-2:3 > gpurun_out/r2h_spmm_2d.jsonl 2> gpurun_out/r2h.err
python tools/bench_spmm.py --shape 256,256,256 --configs 0:64:2:3,1:64:0:3,1:64:-2:3,1:32:-2:3,1:64:2:3 > gpurun_out/r2h_spmm_3d.jsonl 2>> gpurun_out/r2h.err
cat gpurun_out/r2h_spmm_2d.jsonl gpurun_out/r2h_spmm_3d.jsonl
out=gpurun_out/r2h_instep.jsonl
: > $out
for cfg in "0 64 2 3 c2" "1 64 -2 3 c2" "1 64 -3 3 c2" "1 64 0 3 c2-3d" "1 64 -2 3 c2-3d" "0 64 0 3 c2-3d"; do
  set -- $cfg
  MF_SPMM_STRIP=$1 MF_SPMM_STRIP_ROWS=$2 MF_SPMM_STRIP_PFD=$3 MF_SPMM_STRIP_MINB=$4 \
    python bench.py --workload $5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2h.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'cfg':'$cfg','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err']}))" >> $out
done
cat $out

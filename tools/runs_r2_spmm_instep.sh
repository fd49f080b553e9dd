#!/bin/bash
# in-step cost of the CSR product (fused alpha dot, power-capped clocks) for the kernel variants
out=gpurun_out/r2c_instep.jsonl
: > $out
for cfg in "0 64 2 4" "1 64 0 3" "1 32 0 3" "1 64 0 4" "1 64 2 3"; do
  set -- $cfg
  MF_SPMM_STRIP=$1 MF_SPMM_STRIP_ROWS=$2 MF_SPMM_STRIP_PFD=$3 MF_SPMM_STRIP_MINB=$4 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --probes-per-gpu 512 2>>gpurun_out/r2c_instep.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'cfg':'$cfg','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err']}))" >> $out
done
cat $out

#!/bin/bash
# 3-D target, row-group kernel: prefetch distance / block size / window slack sweep (in-step), and the 2-D default
out=gpurun_out/r2w_sweep2.jsonl
: > $out
one() {  # label, env, workload
  env $2 timeout 600 python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2w.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'cfg':'$1','wl':'$3','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate']}))" >> $out
}
one base "MF_SPMM_TMA=0" c2-3d
one mb6 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=6" c2-3d
one mb6pfd5 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=6 MF_SPMM_PFD=5" c2-3d
one mb4pfd5 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=4 MF_SPMM_PFD=5" c2-3d
one mb3pfd5 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=3 MF_SPMM_PFD=5" c2-3d
one mb8pfd5 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=8 MF_SPMM_PFD=5" c2-3d
one mb6pfd7 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=6 MF_SPMM_PFD=7" c2-3d
one mb6pfd5r128 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=6 MF_SPMM_PFD=5 MF_SPMM_ROWS=128" c2-3d
one mb4pfd5r128 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=4 MF_SPMM_PFD=5 MF_SPMM_ROWS=128" c2-3d
one mb2 "MF_SPMM_TMA=0 MF_SPMM_BLOCK_MB=2" c2-3d
one base "MF_SPMM_TMA=0" c2-3d
cat $out

#!/bin/bash
# after making the hybrid 7-diagonal TMA kernel the default for blocked 3-D stencils: parity, the two defaults in the
# step, the row-group kernel beside them, and wider windows for the 2-D kernel
timeout 600 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q 2>&1 | tail -2
out=gpurun_out/r2zd_window.jsonl
: > $out
one() {  # label env workload
  env $2 timeout 600 python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2zd.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'cfg':'$1','wl':'$3','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate']}))" >> $out
}
one default "MF_X=0" c2-3d
one gather "MF_SPMM_TMA=0" c2-3d
one default "MF_X=0" c2-3d
one default "MF_X=0" c2
for w in 49152 98304; do one tma2d_w$w "MF_SPMM_TMA_WINDOW=$w" c2; done
one default "MF_X=0" c2
cat $out

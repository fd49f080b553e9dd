#!/bin/bash
# 3-D target, chunked 7-diagonal TMA kernel (default): block size of the blocked row order (the window follows it)
out=gpurun_out/r2zq_sweep.jsonl
: > $out
one() {  # label, env, workload
  env $2 timeout 600 python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2zq.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'cfg':'$1','wl':'$3','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate']}))" >> $out
}
for w in 8192 6144 12288 16384 8192; do one mb6_w$w "MF_SPMM_TMA_WINDOW=$w" c2-3d; done
cat $out

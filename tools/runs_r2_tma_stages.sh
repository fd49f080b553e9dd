#!/bin/bash
# TMA band kernel: 16 rows x 2 stages against 8 rows x 4 stages (same bytes in flight, finer grain), interleaved
timeout 600 python -m pytest tests/test_gpu_spmm_band.py -q 2>&1 | tail -2
out=gpurun_out/r2z_instep.jsonl
: > $out
for rows in 16 8 16 8; do
  MF_SPMM_TMA_ROWS=$rows timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2z.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'rows':$rows,'value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate']}))" >> $out
done
cat $out

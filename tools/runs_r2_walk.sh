#!/bin/bash
# strip-walk TMA kernel (2-D 5-point bands): parity, then in the step against the chunked TMA kernel (MF_SPMM_WALK=0)
timeout 900 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_configs.py -q -k c2 2>&1 | tail -2
out=gpurun_out/r2ze_walk.jsonl
: > $out
for w in 1 0 1 0; do
  MF_SPMM_WALK=$w timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2ze.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'walk':$w,'value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate'],'rel_err':d['result']['rel_err']}))" >> $out
done
cat $out
tail -3 gpurun_out/r2ze.err

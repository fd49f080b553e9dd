#!/bin/bash
# the TMA-staged band kernel: parity, stand-alone and in-step timings against the other two kernels
timeout 600 python -m pytest tests/test_gpu_spmm_band.py -q 2>&1 | tail -15 > gpurun_out/r2l_tests.log; tail -3 gpurun_out/r2l_tests.log
timeout 300 python tools/bench_spmm.py --configs 0:64:2:3,2:64:0:3,1:64:0:3 > gpurun_out/r2l_spmm_2d.jsonl 2> gpurun_out/r2l.err; cat gpurun_out/r2l_spmm_2d.jsonl
out=gpurun_out/r2l_instep.jsonl
: > $out
for tma in 0 1 0 1; do
  MF_SPMM_TMA=$tma timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2l.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'tma':$tma,'value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err']}))" >> $out
done
cat $out

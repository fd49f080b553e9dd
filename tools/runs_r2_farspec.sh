#!/bin/bash
# producer with the branch-free match for rows that miss entries (in-tree library) against the serial match (ab/lib_head.so)
# library) against all of them after it (ab/lib_head.so), interleaved on the 3-D target
timeout 600 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q 2>&1 | tail -2
out=gpurun_out/r2zo_farspec.jsonl
: > $out
H=$PWD/ab/lib_head.so
for v in new head new head; do
  lib=""; [ $v = head ] && lib=$H
  MF_LIB_PATH=$lib timeout 600 python bench.py --workload c2-3d --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2zo.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$v','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate']}))" >> $out
done
cat $out

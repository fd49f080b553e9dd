#!/bin/bash
# barrier-free TMA band kernel (v3, the in-tree library) against the per-chunk-barrier one (v2, ab/lib_v2.so) and the
# gather kernel: parity first, then interleaved in-step timings (the pods differ by a few percent)
timeout 600 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q 2>&1 | tail -3
out=gpurun_out/r2y_instep.jsonl
: > $out
V2=$PWD/ab/lib_v2.so
for cfg in "v4 1 16" "v2 1 16" "v4 1 8" "v2 1 16" "v4 1 16"; do
  set -- $cfg
  lib=""; [ $1 = v2 ] && lib=$V2
  MF_LIB_PATH=$lib MF_SPMM_TMA=$2 MF_SPMM_TMA_ROWS=$3 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2y.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$1','tma':$2,'rows':$3,'value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err'],'logdet':d['result']['logdet_estimate']}))" >> $out
done
cat $out
tail -5 gpurun_out/r2y.err

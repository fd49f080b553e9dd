#!/bin/bash
# hybrid 7-diagonal TMA band kernel (five staged diagonals + two gathered) on the 3-D target workload, against the
# row-group kernel; the 2-D default beside it (same code path, dense coefficients): parity + in-step timings
timeout 600 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q 2>&1 | tail -3
out=gpurun_out/r2za_instep.jsonl
: > $out
for cfg in "0 16 c2-3d" "2 16 c2-3d" "0 16 c2-3d" "2 16 c2-3d"; do
  set -- $cfg
  MF_SPMM_TMA=$1 MF_SPMM_TMA_ROWS=$2 timeout 600 python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2za.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'wl':'$3','tma':$1,'rows':$2,'value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err'],'logdet':d['result']['logdet_estimate']}))" >> $out
done
cat $out
tail -3 gpurun_out/r2za.err

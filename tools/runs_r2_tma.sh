#!/bin/bash
# the TMA-staged band kernel (dedicated producer warp): parity, in-step timings for 8 / 16 / 32 rows per chunk against
# the gather kernel (interleaved: the pods differ by a few percent), ncu of the best variant
timeout 600 python -m pytest tests/test_gpu_spmm_band.py -q 2>&1 | tail -3
out=gpurun_out/r2n_instep.jsonl
: > $out
for cfg in "0 16" "1 16" "1 8" "1 32" "0 16" "1 16"; do
  set -- $cfg
  MF_SPMM_TMA=$1 MF_SPMM_TMA_ROWS=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2n.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'tma':$1,'rows':$2,'value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err']}))" >> $out
done
cat $out
B="--steps 1 --warmup 1 --profile --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 256"
T=/tmp/ncu_r2
mkdir -p $T
MF_SPMM_TMA=1 timeout 600 ncu --set full --clock-control none -k regex:spmm -s 6 -c 1 -o $T/spmm_2d_tma -f python bench.py $B > /dev/null 2>> gpurun_out/r2n.err
python tools/ncu_summary.py wide $T/spmm_2d_tma.ncu-rep > gpurun_out/r2n_spmm_2d_tma.txt 2>> gpurun_out/r2n.err
grep -E "gpu__time_duration.sum|l1tex__data_pipe_lsu_wavefronts.avg.pct|smsp__inst_executed.sum |dram__bytes_read.sum |smsp__average_warps_issue_stalled_(barrier|long|short|wait|mio|sleeping|selected|not_sel).*ratio|sm__warps_active.avg.pct" gpurun_out/r2n_spmm_2d_tma.txt | cut -c1-140

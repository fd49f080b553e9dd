#!/bin/bash
# in-step numbers after the blocked-order change; C4 on one GPU with / without per-launch event bracketing
out=gpurun_out/r2i_instep.jsonl
: > $out
for wl in c2 c2-3d; do
  python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2i.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'wl':'$wl','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'update_ms':d['kernels']['lanczos_update']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'rel_err':d['result']['rel_err']}))" >> $out
done
cat $out
python tools/bench_c4.py --steps 2 --warmup 1 > gpurun_out/r2i_c4_1gpu_timed.json 2>>gpurun_out/r2i.err
python tools/bench_c4.py --steps 2 --warmup 1 --no-kernel-timing > gpurun_out/r2i_c4_1gpu.json 2>>gpurun_out/r2i.err
python -c "
import json
for f in ('r2i_c4_1gpu_timed','r2i_c4_1gpu'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_decomposition'], d['frac_of_hbm_peak'], d['per_launch_event_bracketing'], d['gpu_launches_per_decomposition'])
"

#!/bin/bash
# strip-walk TMA kernels: parity, then the 3-D target in the step against the chunked hybrid kernel (MF_SPMM_WALK=0)
# and the row-group kernel (MF_SPMM_TMA=0); the 2-D default beside it
timeout 900 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q -x 2>&1 | tail -4
out=gpurun_out/r2zh_walk.jsonl
: > $out
one() {  # label env workload
  env $2 timeout 600 python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>>gpurun_out/r2zh.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'cfg':'$1','wl':'$3','value':d['value'],'spmm_ms':d['kernels']['spmm_csr']['ms_per_launch'],'step_frac':d['step_roofline']['frac_of_peak'],'clk':d['clocks']['sm_mhz'],'logdet':d['result']['logdet_estimate']}))" >> $out
}
one walk "MF_X=0" c2-3d
one chunked "MF_SPMM_WALK=0" c2-3d
one walk_slack1 "MF_SPMM_WALK_SLACK=1" c2-3d
one walk_slack8 "MF_SPMM_WALK_SLACK=8" c2-3d
one walk "MF_X=0" c2-3d
cat $out
tail -3 gpurun_out/r2zh.err

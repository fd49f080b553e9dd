#!/bin/bash
# A/B of a kernel change on the 3-D target: the in-tree library against ab/lib_head.so, the library built from the
# commit before the change (git stash; python -c 'import __graft_entry__ as g; g.build()'; cp matfree_b200/_lib/libmatfree_b200.so ab/lib_head.so;
# git stash pop; rebuild) and selected with MF_LIB_PATH; interleaved because the pods differ by a few percent.
# Used for: the +-plane gathers issued before the stage wait (profiles/r2zn_farspec.jsonl) and the producer's
# branch-free match for rows that miss entries (profiles/r2zo_match.jsonl).
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

#!/bin/bash
# quick regression check of the CSR product kernels: parity tests, then the 2-D and 3-D defaults in the step
timeout 600 python -m pytest tests/test_gpu_spmm_band.py tests/test_gpu_parity.py -q 2>&1 | tail -2
for wl in c2 c2-3d c2 c2-3d; do
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --probes-per-gpu 512 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['value']), round(d['kernels']['spmm_csr']['ms_per_launch'],3), round(d['step_roofline']['frac_of_peak'],4), d['clocks']['sm_mhz'], d['result']['logdet_estimate'])"
done

"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): probe sharding over NCCL gives the
same per-probe values and estimate as one GPU."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MF_ROOT"])
import matfree_b200 as m
from matfree_b200 import workloads
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape, P, k = (96, 96), 300, 12
n = shape[0] * shape[1]
ip, ix, d = workloads.laplacian_csr(shape, shift=1.0, device=f"cuda:{local}")
op = m.ops.csr(ip, ix, d)
key = m.prng.prng_key(7)
sampler = m.stochtrace.sampler_signs(np.broadcast_to(np.float32(1), (n,)), num=P)
integrand = m.funm.monte_carlo_funm_sym_logdet(m.decomp.tridiag_sym(k, reortho="none"))
est = m.stochtrace.estimator_monte_carlo_mean_and_sem(integrand, sampler)
plain = m.stochtrace.estimator_monte_carlo(integrand, sampler)
single = plain.per_probe(op, key, tile=64)          # all probes on this GPU
mean1, sem1 = est(op, key)
with m.stochtrace.probe_sharding():
    sharded = plain.per_probe(op, key, tile=64)     # my shard, all-gathered
    mean2, sem2 = est(op, key)
assert sharded.shape == single.shape == (P,)
assert torch.equal(sharded, single), (sharded - single).abs().max()
assert float(mean1) == float(mean2) and float(sem1) == float(sem2)
# Hutchinson trace, sharded
tr = m.stochtrace.estimator_monte_carlo(m.stochtrace.monte_carlo_trace(), sampler)
t1 = float(tr(op, key))
with m.stochtrace.probe_sharding():
    t2 = float(tr(op, key))
assert t1 == t2
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_probe_sharding_nccl_matches_single_gpu(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MF_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
        env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world


_ROW_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MF_ROOT"])
import matfree_b200 as m
from matfree_b200 import workloads, _rowshard
from oracle import prng as oprng
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
shape = (24, 16, 16)
n = int(np.prod(shape)); plane = shape[1] * shape[2]
k = 20
dev = f"cuda:{local}"
r0, r1 = _rowshard.slab_range(n, world, rank, align=plane)
ip, ix, d = workloads.laplacian_csr_rows(shape, r0, r1, shift=1.0, device=dev)
sop = m.ops.csr_row_sharded(ip, ix, d, n, r0)
assert sop.plan.lo == (plane if rank > 0 else 0) and sop.plan.hi == (plane if rank < world - 1 else 0)
ipf, ixf, df = workloads.laplacian_csr(shape, shift=1.0, device=dev)
op = m.ops.csr(ipf, ixf, df)
v = oprng.rademacher(oprng.prng_key(1), (1, n), np.float32)[0]   # probe 0 of PRNGKey(1)
for reortho in ("full", "none"):
    tri = m.decomp.tridiag_sym(k, reortho=reortho, materialize=False)
    Q1, (d1, e1), res1, c1 = tri(op, v)                 # whole operator on this GPU
    Q2, (d2, e2), res2, c2 = tri(sop, v[r0:r1])         # my slab of the row-sharded run
    assert np.allclose(d2.cpu(), d1.cpu(), rtol=1e-5, atol=1e-5), reortho
    assert np.allclose(e2.cpu(), e1.cpu(), rtol=1e-5, atol=1e-5), reortho
    assert np.allclose(float(c2), float(c1), rtol=1e-6)
    if reortho == "full":
        assert np.allclose(Q2.cpu(), Q1[:, r0:r1].cpu(), atol=1e-4)
        assert np.allclose(res2.cpu(), res1[r0:r1].cpu(), atol=1e-3)
# sharded matvec == slab of the full matvec, bit for bit (same kernel, same row order)
w1 = op(v)[r0:r1]; w2 = sop(v[r0:r1])
assert torch.equal(w1, w2)
# the peer-memory drivers (mf_lanczos_sharded: halo pushed by stores over NVLink, all-reduce fused
# into the reducing kernels) against the NCCL route (Python step loop), on a block of 8 vectors
assert _rowshard._use_peer_memory(None)
V = torch.as_tensor(oprng.normal(oprng.prng_key(9), (n, 8), np.float32)).to(dev)
Vloc = V[r0:r1].contiguous()
for reortho in ("full", "none"):
    a1, b1, l1, Q1, res1 = m.decomp.lanczos_blocked(op, V, k, reortho, want_Q=True, want_residual=True)
    a2, b2, l2, Q2, res2 = m.decomp.lanczos_blocked(sop, Vloc, k, reortho, want_Q=True, want_residual=True)
    sop._comm.check()   # no in-kernel wait timed out
    os.environ["MF_ROWSHARD_NCCL"] = "1"
    a3, b3, l3, Q3, res3 = m.decomp.lanczos_blocked(sop, Vloc, k, reortho, want_Q=True, want_residual=True)
    del os.environ["MF_ROWSHARD_NCCL"]
    for x2, x1, x3 in ((a2, a1, a3), (b2, b1, b3), (l2, l1, l3)):
        assert np.allclose(x2.cpu(), x1.cpu(), rtol=2e-5, atol=2e-5), reortho
        assert np.allclose(x2.cpu(), x3.cpu(), rtol=2e-5, atol=2e-5), reortho
    assert np.allclose(Q2.cpu(), Q1[:, r0:r1].cpu(), atol=2e-4), reortho
    assert np.allclose(res2.cpu(), res1[r0:r1].cpu(), atol=2e-3), reortho
    # the fused all-reduce adds the ranks' sums in rank order on every rank: identical bits
    mine = torch.cat([a2.flatten(), b2.flatten(), l2.flatten()])
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    assert all(torch.equal(t, allv[0]) for t in allv), reortho
    # and is reproducible run to run
    a4, b4, l4, _, _ = m.decomp.lanczos_blocked(sop, Vloc, k, reortho, want_Q=False, want_residual=False)
    assert torch.equal(a4, a2) and torch.equal(b4, b2), reortho
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_row_sharded_tridiag_nccl_matches_single_gpu(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "row_worker.py"
    script.write_text(_ROW_WORKER)
    env = dict(os.environ, MF_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
        env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world

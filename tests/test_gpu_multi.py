"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): probe sharding and row sharding over the
ranks of an NCCL group give the same results as one GPU.  The checks themselves live in
`matfree_b200/_multicheck.py`, which `bench.py` also runs on its live process group at N > 1 so
that the driver's scaling run carries the same evidence in its JSON lines."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["MF_ROOT"])
from matfree_b200 import _multicheck
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = getattr(_multicheck, os.environ["MF_CHECK"])(dev)
bad = {k: v for k, v in res.items() if isinstance(v, bool) and not v}
assert not bad, (rank, res)
dist.barrier()
dist.destroy_process_group()
print("RANK_OK", rank, json.dumps(res))
"""


def _run(tmp_path, check, port):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MF_ROOT=ROOT, MF_CHECK=check)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
        env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("RANK_OK") == world


def test_probe_sharding_nccl_matches_single_gpu(tmp_path):
    _run(tmp_path, "probe_sharding", 29541)


def test_row_sharded_tridiag_nccl_matches_single_gpu(tmp_path):
    _run(tmp_path, "row_sharding", 29543)

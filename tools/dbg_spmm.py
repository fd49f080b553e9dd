import sys
import numpy as np, torch
sys.path.insert(0, ".")
import ctypes, os
from matfree_b200 import _lib
if os.environ.get("MF_LIB_PATH"):
    probe = ctypes.CDLL(os.environ["MF_LIB_PATH"])
    for name in list(_lib.SIGNATURES):
        if not hasattr(probe, name):
            del _lib.SIGNATURES[name]
import matfree_b200 as m
from oracle import prng as oprng
from matfree_b200 import workloads
import scipy.sparse as sp
for dtype in (np.float32, np.float64):
    ip, ix, d = workloads.laplacian_csr((13, 17), shift=0.5, dtype=np.dtype(dtype).name)
    n = 221
    A = sp.csr_matrix((d.numpy(), ix.numpy(), ip.numpy()), shape=(n, n))
    op = m.ops.csr_from_scipy(A)
    for P in (1, 3, 64, 300):
        V = oprng.normal(oprng.prng_key(2), (P, n), dtype)
        got = op.matmat(V).cpu().numpy()
        want = (A @ V.T).T
        print(np.dtype(dtype).name, P, op.max_row_nnz, float(np.abs(got - want).max()))
# direct blocked call (no layout helpers)
import ctypes
from matfree_b200 import _device
lib = _lib.load()
ip, ix, d = workloads.laplacian_csr((13, 17), shift=0.5)
A = sp.csr_matrix((d.numpy(), ix.numpy(), ip.numpy()), shape=(221, 221))
op = m.ops.csr(ip, ix, d)
for ld in (1, 4, 16, 64, 256):
    X = torch.randn(221, ld, device="cuda")
    W = torch.full((221, ld), 7.0, device="cuda")
    rc = lib.mf_matmat_csr(op.indptr.data_ptr(), op.indices.data_ptr(), op.data.data_ptr(), 221, op.nnz, 0,
                           X.data_ptr(), W.data_ptr(), ld, _device.stream())
    torch.cuda.synchronize()
    want = A @ X.cpu().numpy()
    print("direct", ld, rc, float(np.abs(W.cpu().numpy() - want).max()), float((W == 7.0).float().mean()))

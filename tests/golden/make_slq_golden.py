#!/usr/bin/env python
"""Generate tests/golden/slq_golden.json: small SLQ cases evaluated by the ORACLE (oracle/ref.py,
the NumPy restatement of matfree's arithmetic; the reference itself cannot run here -- no JAX --
so these vectors pin the oracle against drift, they are not outputs of a live JAX run).

    python tests/golden/make_slq_golden.py        # rewrites the JSON next to this script

Cases: a 2-D 5-point Laplacian 12 x 11 + 0.75 I (CSR), keys PRNGKey(11) / PRNGKey(3), fp32 and fp64:
per-probe SLQ log-determinant quadratic forms (reortho none / full, depth 9), Ritz values of probe 0,
Hutchinson trace samples, diagonal estimate, Golub-Kahan B of a 20 x 12 matrix and the product
log-determinant quadratic forms.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import prng, ref  # noqa: E402


def laplacian_dense(shape, shift):
    ny, nx = shape
    n = ny * nx
    A = np.zeros((n, n))
    for i in range(ny):
        for j in range(nx):
            r = i * nx + j
            A[r, r] = 4.0 + shift
            if i > 0:
                A[r, r - nx] = -1.0
            if i < ny - 1:
                A[r, r + nx] = -1.0
            if j > 0:
                A[r, r - 1] = -1.0
            if j < nx - 1:
                A[r, r + 1] = -1.0
    return A


def build():
    out = {"about": "oracle-generated (oracle/ref.py); see make_slq_golden.py", "cases": {}}
    shape, shift, P, k = (12, 11), 0.75, 6, 9
    n = shape[0] * shape[1]
    A64 = laplacian_dense(shape, shift)
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        A = A64.astype(dt)
        mm = lambda X: (A @ X.T).T  # noqa: E731
        V = prng.rademacher(prng.prng_key(11), (P, n), dt)
        case = {"shape": list(shape), "shift": shift, "num_probes": P, "depth": k, "key": 11}
        for reortho in ("none", "full"):
            q, theta = ref.slq_batched(mm, V, k, reortho=reortho)
            case[f"quad_{reortho}"] = [float(x) for x in q]
            case[f"ritz_probe0_{reortho}"] = [float(x) for x in theta[0]]
        case["trace_samples"] = [float(v @ (A @ v)) for v in V]
        Vn = prng.normal(prng.prng_key(11), (P, n), dt)
        case["diagonal_mean_normal"] = [float(x) for x in np.mean(Vn * (A @ Vn.T).T, axis=0)]
        case["normal_probe0_head"] = [float(x) for x in Vn[0, :8]]
        # Golub-Kahan on a 20 x 12 matrix with singular values 1..4
        B = ref.asymmetric_matrix_from_singular_values(np.linspace(1.0, 4.0, 12), nrows=20, ncols=12).astype(dt)
        Vb = prng.rademacher(prng.prng_key(3), (4, 12), dt)
        integ = ref.monte_carlo_funm_product_logdet(ref.bidiag(7))
        case["product_logdet_quad"] = [float(integ(B.astype(np.float64), v.astype(np.float64))) for v in Vb]
        (_, _), Bd, _, _ = ref.bidiag(7)(B.astype(np.float64), Vb[0].astype(np.float64))
        case["bidiag_diag_probe0"] = [float(x) for x in np.diag(Bd)]
        case["bidiag_offdiag_probe0"] = [float(x) for x in np.diag(Bd, 1)]
        out["cases"][name] = case
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "slq_golden.json")
    with open(path, "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote", path)

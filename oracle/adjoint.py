"""Oracle restatement of matfree's Lanczos / Arnoldi ADJOINTS in NumPy.  TEST INFRASTRUCTURE.

`matfree/decomp.py:184-217,295-348` (`_tridiag_adjoint`, the custom VJP of
`tridiag_sym(reortho="none")`) and `:398-423,480-600` (`_hessenberg_adjoint`, the custom VJP of
`hessenberg` and hence of `tridiag_sym(reortho="full")`), Kraemer et al. (2024), "Gradients of
functions of large matrices".  The reference takes parameter gradients with `jax.vjp` of the user
matvec; here the matvec is ``matvec(v, A) = A @ v`` with the dense matrix ``A`` as its parameter,
for which those VJPs are outer products (cited per line).  Operation order follows the reference.

Pinned by central finite differences of the forward passes in `oracle/ref.py` (the reference's
own test compares with JAX autodiff of the forward pass: `tests/test_decomp/
test_tridiag_sym_adjoint.py:7-49`, `test_hessenberg_adjoint.py:5-38`), see `tests/test_oracle_adjoint.py`.
"""

from __future__ import annotations

import numpy as np


def tridiag_forward_cache(A, vec, k):
    """`_tridiag_forward` (`decomp.py:220-292`) for ``matvec(v) = A @ v``: the values the custom
    VJP caches -- ``xs (k+1, n)`` (all Lanczos vectors incl. the last), ``alphas (k,)``,
    ``betas (k,)`` (incl. the last) -- and ``|vec|``."""
    from oracle import ref

    (Q, (diags, offdiags)), (q, b), _ = ref._tridiag_forward(lambda x: A @ x, k, vec)
    xs = np.concatenate([Q, q[None]])
    betas = np.concatenate([offdiags, np.asarray(b)[None]])
    return xs, diags, betas, np.linalg.norm(vec)


def tridiag_adjoint(A, *, initvec_norm, alphas, betas, xs, dalphas, dbetas, dxs):
    """`_tridiag_adjoint` + `_tridiag_adjoint_step` (`decomp.py:295-348`).

    Inputs as in `estimate_bwd` (`:189-211`): ``xs, dxs (k+1, n)``, ``alphas, dalphas (k,)``,
    ``betas, dbetas (k,)`` (last entries = the residual's).  Returns ``(grad_initvec (n,),
    grad_A (n, n))`` and the adjoint states ``(lambdas (k, n), mus, nus)``."""
    k = len(alphas)
    xi = -dxs[-1]                                   # :316 init_val
    lam_plus = np.zeros_like(dxs[-1])
    grad = np.zeros_like(A)
    lambdas, mus, nus = [], [], []
    for i in reversed(range(k)):                    # scan(..., reverse=True), :317-319
        dx, da, db = dxs[i], dalphas[i], dbetas[i]
        xplus, x = xs[i + 1], xs[i]
        a, b = alphas[i], betas[i]
        xi = xi / b                                 # :339
        mu = db - lam_plus @ x + xplus @ xi         # :340
        nu = da + x @ xi                            # :341
        lam = -xi + mu * xplus + nu * x             # :342
        matvec_lambda = A @ lam                     # :345
        grad += np.outer(x, lam)                    # :345-346: vjp of p -> p @ lam at cotangent x
        xi = -dx - matvec_lambda + a * lam + b * lam_plus - b * nu * xplus   # :349
        lam_plus = lam
        lambdas.append(lam), mus.append(mu), nus.append(nu)
    lambda_1 = xi                                   # :317 the carry's second slot after the scan
    grad_initvec = ((lambda_1 @ xs[0]) * xs[0] - lambda_1) / initvec_norm   # :324
    return (grad_initvec, grad), (np.array(lambdas[::-1]), np.array(mus[::-1]), np.array(nus[::-1]))


def _extract_diag(x, offset=0):
    # decomp.py:603-605
    return np.diag(np.diagonal(x, offset), offset)


def hessenberg_adjoint(A, *, Q, H, r, c, dQ, dH, dr, dc, reortho):
    """`_hessenberg_adjoint` + `_hessenberg_adjoint_step` (`decomp.py:480-600`).

    ``Q, dQ (n, k)`` (the forward pass's layout, before `estimate` transposes it), ``H, dH
    (k, k)``, ``r, dr (n,)``, ``c, dc`` scalars.  Returns ``(dv (n,), dA (n, n))``."""
    n, k = Q.shape
    if k == 0:
        raise ValueError("Custom Hessenberg-adjoints are not implemented for num_matvecs = 0.")  # :483-486

    def lower(m):
        m_tril = np.tril(m)
        return m_tril - 0.5 * _extract_diag(m_tril)

    eye = np.eye(k, dtype=Q.dtype)
    e_1, e_K = eye[0], eye[-1]
    lower_mask = lower(np.ones((k, k), dtype=Q.dtype))
    gamma = dH @ e_K - Q.T @ dr                     # :498
    lambda_k = dr + Q @ gamma                       # :499
    Lambda = np.zeros_like(Q)
    Gamma = np.zeros((k, k), dtype=Q.dtype)
    dp = np.zeros_like(A)
    Pi_xi = dQ.T + np.outer(gamma, r)               # :505
    Pi_gamma = -dc * c * np.outer(e_1, e_1) + H @ dH.T - (dQ.T @ Q)   # :506
    reortho_mask = np.tril(np.ones((k, k), dtype=Q.dtype), 1)         # :509
    beta_minuses = np.concatenate([np.ones((1,), dtype=Q.dtype), np.diagonal(H, -1)])
    alphas = np.diagonal(H)
    beta_pluses = H - _extract_diag(H) - _extract_diag(H, -1)
    for idx in reversed(range(k)):                  # scan(..., reverse=True), :538
        beta_minus, alpha, beta_plus = beta_minuses[idx], alphas[idx], beta_pluses[idx]
        q = Q[:, idx]
        if reortho == "full":                       # :576-585
            mask = reortho_mask[idx]
            Q_masked = mask[None, :] * Q
            rhs_masked = mask * dH[:, idx]
            lambda_k = lambda_k - Q_masked @ (Q_masked.T @ lambda_k) + Q_masked @ rhs_masked
        vecmat_lambda = A.T @ lambda_k              # :588-589 (vjp wrt the vector)
        dp = dp + np.outer(lambda_k, q)             # :589-590 (vjp wrt the parameter)
        tmp = lower_mask[idx] * (Pi_gamma[idx] - vecmat_lambda @ Q)   # :593
        Gamma[idx, :] = tmp
        Lambda[:, idx] = lambda_k                   # :597
        xi = Pi_xi[idx] + (Gamma + Gamma.T)[idx, :] @ Q.T             # :598
        lambda_k = xi - (alpha * lambda_k - vecmat_lambda) - beta_plus @ Lambda.T   # :599
        lambda_k = lambda_k / beta_minus            # :600
    dv = lambda_k * c                               # :543
    return dv, dp

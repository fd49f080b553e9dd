"""Oracle restatement of matfree's SLQ path in NumPy.  TEST INFRASTRUCTURE.

Every function cites the reference lines it follows (paths relative to
`/root/reference/`).  Operation ORDER follows the reference exactly; the
floating-point type is the dtype of the input vector (the reference computes
in fp32 unless `jax_enable_x64` is set).

Two flavours of every decomposition are given: a single-vector one that reads
like the reference (what `jax.vmap` maps over), and a probe-batched one
(`*_batched`, rows = probes) that performs the same operations for a whole
`(P, n)` block at once -- that is what XLA executes after `vmap`
(`matfree/stochtrace.py:49`) and what the CPU baseline times.
"""

from __future__ import annotations

from typing import Callable, NamedTuple

import numpy as np

from oracle import prng as _prng


class DecompResult(NamedTuple):
    """`matfree/decomp.py:15-27` (_DecompResult)."""

    Q_tall: np.ndarray
    J_small: object
    residual: np.ndarray
    init_length_inv: object


def _error_num_matvecs(num, maxval, minval):
    # matfree/decomp.py:753-756
    msg1 = f"Parameter 'num_matvecs'={num} exceeds the acceptable range. "
    msg2 = f"Expected: {minval} <= num_matvecs <= {maxval}."
    return msg1 + msg2


def todense_tridiag_sym(diag, off_diag):
    # matfree/decomp.py:148-152
    return np.diag(diag) + np.diag(off_diag, -1) + np.diag(off_diag, 1)


# --------------------------------------------------------------------------
# tridiag_sym, reortho="none"  (three-term Lanczos)
# --------------------------------------------------------------------------


def _tridiag_forward(matvec, num_matvecs, vec):
    """`matfree/decomp.py:220-254` with `_tridiag_fwd_init :257-268` and
    `_tridiag_fwd_step_apply :286-292`."""
    dt = vec.dtype
    k = num_matvecs
    vectors = np.zeros((k + 1, len(vec)), dtype=dt)
    offdiags = np.zeros((k,), dtype=dt)
    diags = np.zeros((k,), dtype=dt)

    v0 = vec / np.linalg.norm(vec).astype(dt)  # :227
    vectors[0] = v0

    # init step :257-268
    Av = matvec(v0)
    a = v0 @ Av
    r = Av - a * v0
    b = np.linalg.norm(r).astype(dt)
    v1 = r / b
    inv_len = (1 / np.linalg.norm(vec)).astype(dt)
    if k == 0:  # :233-236
        return (vectors[:-1], (diags, offdiags[:-1])), (v1, b), inv_len

    vectors[1] = v1
    offdiags[0] = b
    diags[0] = a
    vprev = v0
    for i in range(1, k):  # fori_loop :247-249
        Av = matvec(v1)
        a = v1 @ Av  # :288
        r = Av - a * v1 - b * vprev  # :289 (left to right)
        b = np.linalg.norm(r).astype(dt)  # :290
        x = r / b  # :291
        vprev, v1 = v1, x
        vectors[i + 1] = v1
        offdiags[i] = b
        diags[i] = a
    return (vectors[:-1], (diags, offdiags[:-1])), (v1, b), inv_len


def _tridiag_reortho_none(num_matvecs, *, materialize):
    # matfree/decomp.py:155-182
    def estimate(matvec, vec):
        vec = np.asarray(vec)
        if num_matvecs < 0 or num_matvecs > len(vec):
            raise ValueError(_error_num_matvecs(num_matvecs, maxval=len(vec), minval=0))
        (Q, H), (q, b), _ = _tridiag_forward(matvec, num_matvecs, vec)
        v_flat = b * q  # :167
        if materialize:
            H = todense_tridiag_sym(*H)
        length = np.linalg.norm(vec).astype(vec.dtype)
        return DecompResult(Q, H, v_flat, vec.dtype.type(1.0) / length)

    return estimate


# --------------------------------------------------------------------------
# tridiag_sym, reortho="full"  (Arnoldi + CGS2, T = (H + H^T)/2)
# --------------------------------------------------------------------------


def _hessenberg_forward(matvec, num_matvecs, v, reortho="full"):
    """`matfree/decomp.py:426-477`."""
    if num_matvecs < 0 or num_matvecs > len(v):
        raise ValueError(_error_num_matvecs(num_matvecs, maxval=len(v), minval=0))
    dt = v.dtype
    n, k = len(v), num_matvecs
    Q = np.zeros((n, k), dtype=dt)
    H = np.zeros((k, k), dtype=dt)
    initlength = np.sqrt(np.inner(v, v)).astype(dt)  # :435
    length = initlength
    v = v.copy()
    for i in range(k):  # :448, body :454-477
        v = v / length  # :456
        Q[:, i] = v  # :457
        v = matvec(v)  # :460
        h = Q.T @ v  # :463
        v = v - Q @ h  # :464
        if reortho != "none":
            v = v - Q @ (Q.T @ v)  # :467-468 (h NOT updated)
        length = np.sqrt(np.inner(v, v)).astype(dt)  # :471
        if i + 1 < k:  # :474 (out-of-bounds write is dropped by JAX)
            h[i + 1] = length
        H[:, i] = h  # :475
    return Q, H, v, dt.type(1) / initlength


def _tridiag_reortho_full(num_matvecs, *, materialize):
    # matfree/decomp.py:125-145
    def estimate(matvec, vec):
        vec = np.asarray(vec)
        Q, H, v, c = _hessenberg_forward(matvec, num_matvecs, vec)
        T = vec.dtype.type(0.5) * (H + H.T)  # :133
        diags = np.diagonal(T, 0).copy()
        offdiags = np.diagonal(T, 1).copy()
        matrix = (diags, offdiags)
        if materialize:
            matrix = todense_tridiag_sym(diags, offdiags)
        # :132,142 -- `norm` is hessenberg's init_length_inv = 1/|v| and the result is built with
        # `init_length_inv=1.0 / norm`: the full variant returns the LENGTH, not its inverse
        # (an upstream quirk; `reortho="none"` returns 1/|v|, :177).  Restated as is.
        return DecompResult(Q.T, matrix, v, vec.dtype.type(1.0) / c)  # Q transposed at :388

    return estimate


def tridiag_sym(num_matvecs, /, *, materialize: bool = True, reortho: str = "full",
                custom_vjp: bool = True):
    """`matfree/decomp.py:30-122` (the `custom_vjp` flag only changes gradients)."""
    del custom_vjp
    if reortho == "full":
        return _tridiag_reortho_full(num_matvecs, materialize=materialize)
    if reortho == "none":
        return _tridiag_reortho_none(num_matvecs, materialize=materialize)
    msg = f"reortho={reortho} unsupported. Choose eiter {'full', 'none'}."
    raise ValueError(msg)


# --------------------------------------------------------------------------
# funm: dense function, SLQ integrand, Lanczos action
# --------------------------------------------------------------------------


def dense_funm_sym_eigh(matfun: Callable):
    """`matfree/funm.py:322-335`."""

    def fun(dense_matrix):
        eigvals, eigvecs = np.linalg.eigh(dense_matrix)
        fx = matfun(eigvals)
        return eigvecs @ np.diag(fx) @ eigvecs.T

    return fun


def monte_carlo_funm_sym(dense_funm, tridiag, /):
    """`matfree/funm.py:205-243`."""

    def quadform(matvec, v0):
        v0 = np.asarray(v0)
        length = np.linalg.norm(v0).astype(v0.dtype)  # :228
        v0n = v0 / length  # :229
        _, dense, *_ = tridiag(matvec, v0n)  # :237
        fA = dense_funm(dense)  # :239
        e1 = np.eye(len(fA), dtype=v0.dtype)[0, :]
        return length**2 * np.inner(e1, fA @ e1)  # :241

    return quadform


def monte_carlo_funm_sym_logdet(tridiag, /):
    """`matfree/funm.py:186-202`."""
    return monte_carlo_funm_sym(dense_funm_sym_eigh(np.log), tridiag)


def funm_lanczos_sym(dense_funm, tridiag, /):
    """`matfree/funm.py:114-147`."""

    def estimate(matvec, vec):
        vec = np.asarray(vec)
        length = np.linalg.norm(vec).astype(vec.dtype)
        vecn = vec / length
        Q, matrix, *_ = tridiag(matvec, vecn)
        funm = dense_funm(matrix)
        e1 = np.eye(len(matrix), dtype=vec.dtype)[0, :]
        return length * (Q.T @ (funm @ e1))

    return estimate


# --------------------------------------------------------------------------
# stochtrace: samplers, integrands, estimators
# --------------------------------------------------------------------------


def sampler_signs(n, *, num, dtype=np.float32, mode="partitionable", x64=None):
    """`matfree/stochtrace.py:932-937,957-977` for a flat real vector of length n."""

    def sample(key, p0=0, p1=None):
        p1 = num if p1 is None else p1
        return _prng.rademacher(key, (p1 - p0, n), dtype, mode=mode, x64=x64, offset=p0 * n)

    return sample


def sampler_normal(n, *, num, dtype=np.float32, mode="partitionable"):
    """`matfree/stochtrace.py:927-929,957-964`."""

    def sample(key, p0=0, p1=None):
        p1 = num if p1 is None else p1
        return _prng.normal(key, (p1 - p0, n), dtype, mode=mode, offset=p0 * n)

    return sample


def monte_carlo_trace():
    """`matfree/stochtrace.py:853-865`."""

    def integrand(matvec, v):
        return np.inner(v, matvec(v))

    return integrand


def monte_carlo_diagonal():
    """`matfree/stochtrace.py:836-849` (real dtypes: conj is the identity)."""

    def integrand(matvec, v):
        return v * matvec(v)

    return integrand


def monte_carlo_trace_and_diagonal():
    """`matfree/stochtrace.py:868-883`."""

    def integrand(matvec, v):
        qv = matvec(v)
        return {"trace": np.inner(v, qv), "diagonal": v * qv}

    return integrand


def monte_carlo_rownorms_squared():
    """`matfree/stochtrace.py:886-898`."""

    def integrand(matvec, v):
        qv = matvec(v)
        return qv * qv

    return integrand


def monte_carlo_frobeniusnorm_squared():
    """`matfree/stochtrace.py:901-914`."""

    def integrand(matvec, v):
        x = matvec(v)
        return np.inner(x, x)

    return integrand


def _stack(vals):
    """The sample axis `vmap` adds (`stochtrace.py:49`), for arrays and one-level dicts."""
    if isinstance(vals[0], dict):
        return {k: np.stack([v[k] for v in vals]) for k in vals[0]}
    return np.stack(vals)


def _tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: fn(v) for k, v in tree.items()}
    return fn(tree)


def estimator_monte_carlo(integrand, /, sampler):
    """`matfree/stochtrace.py:7-52`."""

    def estimate(matvec, key):
        samples = sampler(key)
        qs = _stack([integrand(matvec, s) for s in samples])
        return _tree_map(lambda q: np.mean(q, axis=0), qs)

    return estimate


def estimator_monte_carlo_mean_and_sem(integrand, /, sampler):
    """`matfree/stochtrace.py:55-89` (std with ddof=0)."""

    def estimate(matvec, key):
        samples = sampler(key)
        qs = _stack([integrand(matvec, s) for s in samples])
        return (_tree_map(lambda q: np.mean(q, axis=0), qs),
                _tree_map(lambda q: np.std(q, axis=0) / np.sqrt(q.shape[0]), qs))

    return estimate


# --------------------------------------------------------------------------
# probe-batched versions (what XLA runs after vmap); rows of V are probes
# --------------------------------------------------------------------------


def _rowdot(a, b):
    return np.einsum("pn,pn->p", a, b)


def lanczos_none_batched(matmat, V, k):
    """Batched `matfree/decomp.py:220-292`.  `matmat(X)` maps `(P, n) -> (P, n)`
    (row p is `A @ X[p]`).  Returns `(alphas (P,k), betas (P,k), init_length (P,))`;
    `betas[:, :k-1]` are the off-diagonals, `betas[:, k-1]` the residual norm."""
    dt = V.dtype
    P = V.shape[0]
    alphas = np.zeros((P, k), dtype=dt)
    betas = np.zeros((P, k), dtype=dt)
    length = np.sqrt(_rowdot(V, V)).astype(dt)
    v = V / length[:, None]
    vprev = np.zeros_like(v)
    b = np.zeros((P,), dtype=dt)
    for i in range(k):
        w = matmat(v)
        a = _rowdot(v, w).astype(dt)
        r = w - a[:, None] * v - b[:, None] * vprev
        b = np.sqrt(_rowdot(r, r)).astype(dt)
        vprev, v = v, r / b[:, None]
        alphas[:, i] = a
        betas[:, i] = b
    return alphas, betas, length


def lanczos_full_batched(matmat, V, k):
    """Batched `matfree/decomp.py:426-477` + `:133-135`.
    Returns `(diags (P,k), offdiags (P,k-1), init_length (P,))`."""
    dt = V.dtype
    P, n = V.shape
    Q = np.zeros((P, k, n), dtype=dt)
    H = np.zeros((P, k, k), dtype=dt)
    length = np.sqrt(_rowdot(V, V)).astype(dt)
    init = length.copy()
    v = V.copy()
    for i in range(k):
        v = v / length[:, None]
        Q[:, i, :] = v
        v = matmat(v)
        h = np.einsum("pkn,pn->pk", Q, v)
        v = v - np.einsum("pkn,pk->pn", Q, h)
        v = v - np.einsum("pkn,pk->pn", Q, np.einsum("pkn,pn->pk", Q, v))
        length = np.sqrt(_rowdot(v, v)).astype(dt)
        if i + 1 < k:
            h[:, i + 1] = length
        H[:, :, i] = h
    T = dt.type(0.5) * (H + np.swapaxes(H, 1, 2))
    diags = np.diagonal(T, 0, 1, 2).copy()
    offdiags = np.diagonal(T, 1, 1, 2).copy()
    return diags, offdiags, init


def quadrature_batched(diags, offdiags, lengths, matfun=np.log):
    """Batched `matfree/funm.py:239-241,330-333`: ``len^2 * e1^T f(T) e1`` via eigh.
    Also returns the Ritz values (sorted) for parity checks."""
    P, k = diags.shape
    T = np.zeros((P, k, k), dtype=diags.dtype)
    idx = np.arange(k)
    T[:, idx, idx] = diags
    if k > 1:
        T[:, idx[:-1], idx[1:]] = offdiags[:, : k - 1]
        T[:, idx[1:], idx[:-1]] = offdiags[:, : k - 1]
    theta, S = np.linalg.eigh(T)
    q = lengths**2 * np.einsum("pj,pj->p", matfun(theta), S[:, 0, :] ** 2)
    return q.astype(diags.dtype), theta


def slq_batched(matmat, V, k, *, reortho="none", matfun=np.log):
    """Per-probe SLQ quadratic forms for a probe block `V (P, n)`."""
    if reortho == "none":
        a, b, length = lanczos_none_batched(matmat, V, k)
        return quadrature_batched(a, b[:, : k - 1], length, matfun)
    if reortho == "full":
        d, e, length = lanczos_full_batched(matmat, V, k)
        return quadrature_batched(d, e, length, matfun)
    raise ValueError(reortho)


# --------------------------------------------------------------------------
# Golub-Kahan bidiagonalisation and functions of A^T A
# --------------------------------------------------------------------------


def bidiag(num_matvecs, /, materialize=True, reortho="full"):
    """`matfree/decomp.py:608-750`, same operation order; `A` is a dense (nrows, ncols) array
    (the reference gets the vector-matrix product from `jax.vjp` of the matvec)."""

    def estimate(A, v0):
        A = np.asarray(A)
        dt = np.asarray(v0).dtype
        nrows, ncols = A.shape
        k = num_matvecs
        if k > min(nrows, ncols) or k < 0:
            raise ValueError(_error_num_matvecs(k, maxval=min(nrows, ncols), minval=0))
        length = np.linalg.norm(v0).astype(dt)
        v0n = v0 / length                               # :660
        alphas = np.zeros((k,), dt)
        betas = np.zeros((k,), dt)
        Us = np.zeros((k, nrows), dt)
        Vs = np.zeros((k, ncols), dt)
        vk = v0n / np.linalg.norm(v0n).astype(dt)       # :697
        beta = dt.type(0)
        for i in range(k):
            Vs[i] = vk
            betas[i] = beta
            uk = A @ vk - beta * Us[i - 1]              # :703-704
            if reortho == "full":
                uk = uk - Us.T @ (Us @ uk)
                uk = uk - Us.T @ (Us @ uk)
            alpha = np.linalg.norm(uk).astype(dt)
            uk = uk / alpha
            Us[i] = uk
            alphas[i] = alpha
            vk = A.T @ uk - alpha * vk                  # :712-713
            if reortho == "full":
                vk = vk - Vs.T @ (Vs @ vk)
                vk = vk - Vs.T @ (Vs @ vk)
            beta = np.linalg.norm(vk).astype(dt)
            vk = vk / beta
        if materialize:
            J = np.diag(alphas) + np.diag(betas[1:], 1)
        else:
            J = (alphas, betas[1:])
        return (Us, Vs), J, beta * vk, 1 / length

    return estimate


def dense_funm_product_svd(matfun):
    """`matfree/funm.py:305-319`."""

    def dense_funm(matrix):
        _, S, Vt = np.linalg.svd(matrix, full_matrices=False)
        eigvals, eigvecs = S**2, Vt.T
        return eigvecs @ (matfun(eigvals)[:, None] * eigvecs.T)

    return dense_funm


def monte_carlo_funm_product(dense_funm, bidiag_alg, /):
    """`matfree/funm.py:275-302`; the `matvec` argument is the dense matrix itself."""

    def quadform(A, v0):
        length = np.linalg.norm(v0).astype(v0.dtype)
        _, B, *_ = bidiag_alg(A, v0 / length)
        fA = dense_funm(B)
        return length**2 * fA[0, 0]

    return quadform


def monte_carlo_funm_product_logdet(bidiag_alg, /):
    """`matfree/funm.py:246-255`."""
    return monte_carlo_funm_product(dense_funm_product_svd(np.log), bidiag_alg)


def monte_carlo_funm_product_schatten_norm(power, bidiag_alg, /):
    """`matfree/funm.py:258-272`."""
    return monte_carlo_funm_product(dense_funm_product_svd(lambda x: x ** (power / 2)), bidiag_alg)


def hessenberg(num_matvecs, /, *, reortho, reortho_vjp="match"):
    """`matfree/decomp.py:351-477` (forward pass): returns ``(Q (k, n), H, residual, 1/|v|)``.

    The forward pass is called with ``reortho=reortho_vjp`` (`decomp.py:393-396`), whose default
    is the string "match": `_hessenberg_forward_step` only tests ``reortho != "none"`` (`:466`),
    so the second Gram-Schmidt pass runs unless ``reortho_vjp == "none"`` -- also for
    ``reortho="none"``, which only selects the adjoint's re-projection.  Restated as is."""
    if reortho not in ("none", "full"):
        raise TypeError(f"Unexpected input for {reortho}: either of {['none', 'full']} expected.")

    def estimate(matvec, v):
        Q, H, r, c = _hessenberg_forward(matvec, num_matvecs, np.asarray(v), reortho=reortho_vjp)
        return Q.T, H, r, c

    return estimate


def eigh_partial(tridiag_alg):
    """`matfree/eig.py:69-104`: Ritz values / vectors from a tridiagonalisation."""

    def eigh(Av, v0):
        Q, H, *_ = tridiag_alg(Av, v0)
        vals, vecs = np.linalg.eigh(H)
        return vals, vecs.T @ Q

    return eigh


def svd_partial(bidiag_alg):
    """`matfree/eig.py:22-66`."""

    def svd(A, v0):
        (u, v), B, *_ = bidiag_alg(A, v0)
        U, S, Vt = np.linalg.svd(B, full_matrices=False)
        return U.T @ u, S, Vt @ v

    return svd


# --------------------------------------------------------------------------
# matfree/test_util.py restated (fixtures for the parity tests)
# --------------------------------------------------------------------------


def hermitian_matrix_from_eigenvalues(eigvals, key, *, dtype=None, mode="partitionable"):
    """`matfree/test_util.py:6-16` (real case): QR of a normal matrix."""
    eigvals = np.asarray(eigvals)
    (n,) = eigvals.shape
    dtype = eigvals.dtype if dtype is None else np.dtype(dtype)
    X = _prng.normal(key, (n, n), dtype, mode=mode)
    Q, _ = np.linalg.qr(X)
    return ((Q * eigvals) @ Q.T).astype(dtype)


def asymmetric_matrix_from_singular_values(vals, /, nrows, ncols):
    """`matfree/test_util.py:33-38`."""
    A = np.reshape(np.arange(1.0, nrows * ncols + 1.0), (nrows, ncols))
    A /= nrows * ncols
    U, _S, Vt = np.linalg.svd(A, full_matrices=False)
    return U @ np.diag(vals) @ Vt


def assert_allclose(a, b, /, atol=None, rtol=None):
    """`matfree/test_util.py:70-88` (note: the reference passes rtol=atol)."""
    a = np.asarray(a)
    b = np.asarray(b)
    tol = 10 * np.sqrt(np.finfo(a.dtype).eps)
    if tol < 1e-6:
        tol *= 10
    rtol = rtol if rtol is not None else tol
    atol = atol if atol is not None else tol
    assert np.allclose(a, b, atol=atol, rtol=atol), np.max(np.abs(a - b))


def assert_columns_orthonormal(Q, /):
    """`matfree/test_util.py:63-67`."""
    eye_like = Q.T @ Q
    assert_allclose(eye_like, np.eye(len(eye_like), dtype=Q.dtype))

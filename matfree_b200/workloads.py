"""Synthetic operators of the BASELINE.json configurations.

These are workload generators (inputs), not reference functionality: stencil
Laplacians in CSR, the tutorial-1 dense SPD matrix, and closed-form spectra for
checking results at sizes where no dense ground truth is computable.
"""

from __future__ import annotations

import numpy as np


def laplacian_csr(shape, shift=1.0, dtype="float32", device="cpu"):
    """(2d+1)-point Dirichlet Laplacian on a d-dimensional grid plus `shift * I`, as CSR.

    Row index is the row-major flat grid index; diagonal ``2d + shift``, off-diagonals -1;
    columns ascending within a row; int32 indices.  Returns torch tensors
    ``(indptr, indices, data)`` on `device`.
    """
    import torch

    shape = tuple(int(s) for s in shape)
    d = len(shape)
    n = int(np.prod(shape))
    dev = torch.device(device)
    tdt = torch.float64 if str(dtype).endswith("64") else torch.float32
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    strides = [int(np.prod(shape[a + 1:])) for a in range(d)]
    coords = [(idx // strides[a]) % shape[a] for a in range(d)]
    # neighbour offsets in ascending column order: -s0, -s1, ..., 0, ..., +s1, +s0
    cols, masks, vals = [], [], []
    for a in range(d):
        cols.append(idx - strides[a])
        masks.append(coords[a] > 0)
        vals.append(-1.0)
    cols.append(idx)
    masks.append(torch.ones(n, dtype=torch.bool, device=dev))
    vals.append(2.0 * d + shift)
    for a in reversed(range(d)):
        cols.append(idx + strides[a])
        masks.append(coords[a] < shape[a] - 1)
        vals.append(-1.0)
    del coords
    mask = torch.stack(masks, dim=1)  # [n, 2d+1]
    del masks
    counts = mask.sum(dim=1)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, dim=0)
    colmat = torch.stack(cols, dim=1)
    del cols
    indices = colmat[mask].to(torch.int32)
    del colmat
    valrow = torch.tensor(vals, dtype=tdt, device=dev)
    data = valrow.expand(n, -1)[mask].contiguous()
    return indptr.to(torch.int32), indices, data


def laplacian_csr_rows(shape, r0, r1, shift=1.0, dtype="float32", device="cpu"):
    """Rows ``[r0, r1)`` of `laplacian_csr(shape, shift)`: local `indptr` (starting at 0), GLOBAL
    column indices (int64), `data` -- what one rank of a row-sharded run builds, without ever
    materialising the other rows."""
    import torch

    shape = tuple(int(s) for s in shape)
    d = len(shape)
    dev = torch.device(device)
    tdt = torch.float64 if str(dtype).endswith("64") else torch.float32
    idx = torch.arange(int(r0), int(r1), dtype=torch.int64, device=dev)
    nloc = idx.numel()
    strides = [int(np.prod(shape[a + 1:])) for a in range(d)]
    coords = [(idx // strides[a]) % shape[a] for a in range(d)]
    cols, masks, vals = [], [], []
    for a in range(d):
        cols.append(idx - strides[a])
        masks.append(coords[a] > 0)
        vals.append(-1.0)
    cols.append(idx)
    masks.append(torch.ones(nloc, dtype=torch.bool, device=dev))
    vals.append(2.0 * d + shift)
    for a in reversed(range(d)):
        cols.append(idx + strides[a])
        masks.append(coords[a] < shape[a] - 1)
        vals.append(-1.0)
    mask = torch.stack(masks, dim=1)
    indptr = torch.zeros(nloc + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(mask.sum(dim=1), dim=0)
    indices = torch.stack(cols, dim=1)[mask]
    valrow = torch.tensor(vals, dtype=tdt, device=dev)
    data = valrow.expand(nloc, -1)[mask].contiguous()
    return indptr.to(torch.int32), indices, data


def laplacian_eigenvalues(shape, shift=1.0):
    """Closed-form spectrum of `laplacian_csr(shape, shift)` (fp64, unsorted)."""
    lam = np.zeros((1,), dtype=np.float64)
    for m in shape:
        p = np.arange(1, m + 1, dtype=np.float64)
        lam_a = 2.0 - 2.0 * np.cos(p * np.pi / (m + 1))
        lam = (lam[:, None] + lam_a[None, :]).reshape(-1)
    return lam + shift


def laplacian_logdet(shape, shift=1.0):
    """Exact log-determinant of the shifted Laplacian, accumulated axis by axis in fp64."""
    shape = tuple(int(s) for s in shape)
    if len(shape) == 1:
        return float(np.sum(np.log(laplacian_eigenvalues(shape, shift))))
    head = laplacian_eigenvalues(shape[:-1], 0.0)
    p = np.arange(1, shape[-1] + 1, dtype=np.float64)
    last = 2.0 - 2.0 * np.cos(p * np.pi / (shape[-1] + 1)) + shift
    total = 0.0
    for chunk in np.array_split(head, max(1, head.size // 4096)):
        total += float(np.sum(np.log(chunk[:, None] + last[None, :])))
    return total


def powerlaw_laplacian_csr(n, num_edges, gamma=2.3, i0=10.0, seed=5, dtype="float32", device="cpu"):
    """Config 5: graph Laplacian ``L = D - W`` of a Chung-Lu random graph with power-law expected
    degrees, ``w_i ~ (i + i0)^(-1/(gamma-1))``: `num_edges` endpoint pairs
    are drawn i.i.d. from the weight distribution, self-loops and duplicates dropped, the rest
    symmetrised.  Returns ``(indptr, indices, data, d_max)`` as torch tensors on `device` (CSR,
    int32, columns ascending, diagonal = degree, off-diagonals -1).  The stream is torch's
    generator with `seed` (a workload generator, not reference functionality)."""
    import torch

    dev = torch.device(device)
    n = int(n)
    tdt = torch.float64 if str(dtype).endswith("64") else torch.float32
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    w = (torch.arange(n, dtype=torch.float64, device=dev) + float(i0)) ** (-1.0 / (gamma - 1.0))
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    del w
    keys = []
    chunk = 1 << 24
    for e0 in range(0, int(num_edges), chunk):
        m = min(chunk, int(num_edges) - e0)
        src = torch.searchsorted(cdf, torch.rand(m, dtype=torch.float64, device=dev, generator=gen)).clamp_(max=n - 1)
        dst = torch.searchsorted(cdf, torch.rand(m, dtype=torch.float64, device=dev, generator=gen)).clamp_(max=n - 1)
        keep = src != dst
        src, dst = src[keep], dst[keep]
        keys.append(src * n + dst)
        keys.append(dst * n + src)
    del cdf
    keys = torch.unique(torch.cat(keys))           # sorted, duplicates removed
    rows = keys // n
    deg = torch.bincount(rows, minlength=n)
    del rows
    diag = torch.arange(n, dtype=torch.int64, device=dev)
    keys = torch.sort(torch.cat([keys, diag * n + diag])).values
    rows = keys // n
    cols = keys - rows * n
    del keys
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(deg + 1, 0)
    is_diag = rows == cols
    data = torch.where(is_diag, deg[rows].to(tdt), torch.tensor(-1.0, dtype=tdt, device=dev))
    return indptr.to(torch.int32), cols.to(torch.int32), data, int(deg.max())


def tutorial1_dense(n=1000, nrows=1200, dtype=np.float32):
    """Config 1: ``M = A0^T A0 + I`` with ``A0 = reshape(arange(1, 1+nrows*n)) / (nrows*n)``
    (tutorials/1_log_determinants.py:14-21, scaled to n=1000).  Returns (M, A0)."""
    A0 = (np.arange(1.0, 1.0 + nrows * n).reshape(nrows, n) / (nrows * n)).astype(np.float64)
    M = A0.T @ A0 + np.eye(n)
    return M.astype(dtype), A0.astype(dtype)

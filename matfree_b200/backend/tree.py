"""Pytree flattening for device vectors -- mirrors `matfree/backend/tree.py:15-20`
(`jax.flatten_util.ravel_pytree`) for the containers matfree's users pass as vectors: nested
`dict` (keys in sorted order, as JAX flattens them), `list`, `tuple` and `None`, with array
leaves.  Leaves become CUDA tensors; the flat vector is their concatenation in leaf order.  (`device` is for the host-logic tests;
the library itself always places vectors on the current CUDA device.)"""

from __future__ import annotations

import numpy as np

from matfree_b200 import _device


def is_leaf(x) -> bool:
    return not isinstance(x, (dict, list, tuple)) and x is not None


def tree_leaves(tree):
    out = []

    def visit(x):
        if isinstance(x, dict):
            for key in sorted(x):
                visit(x[key])
        elif isinstance(x, (list, tuple)):
            for y in x:
                visit(y)
        elif x is not None:
            out.append(x)

    visit(tree)
    return out


def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {key: tree_map(fn, tree[key]) for key in tree}
    if isinstance(tree, tuple):
        return tuple(tree_map(fn, y) for y in tree)
    if isinstance(tree, list):
        return [tree_map(fn, y) for y in tree]
    if tree is None:
        return None
    return fn(tree)


def _rebuild(tree, leaves_iter):
    if isinstance(tree, dict):
        new = {key: None for key in tree}
        for key in sorted(tree):
            new[key] = _rebuild(tree[key], leaves_iter)
        return new
    if isinstance(tree, tuple):
        return tuple(_rebuild(y, leaves_iter) for y in tree)
    if isinstance(tree, list):
        return [_rebuild(y, leaves_iter) for y in tree]
    if tree is None:
        return None
    return next(leaves_iter)


def leaf_shapes(tree):
    return [tuple(getattr(leaf, "shape", np.shape(leaf))) for leaf in tree_leaves(tree)]


def _as_tensor(leaf, dtype, device):
    import torch

    if device is None:
        return _device.as_device(leaf, dtype)
    t = leaf if isinstance(leaf, torch.Tensor) else torch.as_tensor(np.asarray(leaf))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device).contiguous()


def ravel_pytree(tree, dtype=None, device=None):
    """``(flat, unravel)``: `flat` is a 1-D CUDA tensor, ``unravel(flat_like)`` rebuilds the
    structure (leaf shapes restored); ``unravel.batched(mat)`` does so for every row of a
    ``(k, n)`` tensor, giving leaves with a leading ``k`` axis (the reference's
    ``vmap(unravel)``, `matfree/decomp.py:171,388`)."""
    import torch

    leaves = [_as_tensor(leaf, dtype, device) for leaf in tree_leaves(tree)]
    shapes = [tuple(leaf.shape) for leaf in leaves]
    sizes = [int(np.prod(s)) if s else 1 for s in shapes]
    if leaves:
        common = leaves[0].dtype
        for leaf in leaves[1:]:
            common = torch.promote_types(common, leaf.dtype)
        flat = torch.cat([leaf.reshape(-1).to(common) for leaf in leaves])
    else:
        flat = torch.empty((0,), dtype=dtype or torch.float32, device=device or _device.device())
    trivial = is_leaf(tree) and len(shapes[0]) == 1

    def unravel(vec):
        if trivial:
            return vec
        parts, off = [], 0
        for shape, size in zip(shapes, sizes):
            parts.append(vec[off:off + size].reshape(shape))
            off += size
        return _rebuild(tree, iter(parts))

    def batched(mat):
        if trivial:
            return mat
        parts, off = [], 0
        for shape, size in zip(shapes, sizes):
            parts.append(mat[:, off:off + size].reshape((mat.shape[0],) + shape))
            off += size
        return _rebuild(tree, iter(parts))

    unravel.batched = batched
    unravel.trivial = trivial
    return flat, unravel

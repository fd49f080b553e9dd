"""Device plumbing: torch is used for device memory, streams and NCCL only."""

from __future__ import annotations

import numpy as np


def _torch():
    import torch

    return torch


def device():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("matfree_b200 needs a CUDA device (B200); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def stream() -> int:
    return _torch().cuda.current_stream().cuda_stream


def torch_dtype(dtype):
    torch = _torch()
    if dtype is None:
        return torch.float32
    if isinstance(dtype, torch.dtype):
        if dtype not in (torch.float32, torch.float64):
            raise TypeError(f"dtype {dtype} unsupported (float32 / float64 only)")
        return dtype
    dt = np.dtype(dtype)
    if dt == np.float32:
        return torch.float32
    if dt == np.float64:
        return torch.float64
    raise TypeError(f"dtype {dtype} unsupported (float32 / float64 only)")


def mf_dtype(tdt) -> int:
    return 1 if tdt == _torch().float64 else 0


def as_device(x, dtype=None):
    """Array-like (NumPy / torch / list) -> contiguous CUDA tensor."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device()).contiguous()


def ld_for(num: int, cap: int = 256) -> int:
    """Tile width: the smallest power of two >= num, capped at `cap`."""
    ld = 1
    while ld < num and ld < cap:
        ld *= 2
    return ld


def workspace(nbytes: int):
    torch = _torch()
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device())

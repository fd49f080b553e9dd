"""Probe sharding across ranks (one process per GPU, `torch.distributed`).

Probes are independent and the PRNG is counter-based, so rank r simply evaluates
probes ``[p0, p1)`` of the single-device sample array; the only data-path
collective is one all-gather of the per-probe values (<= 64 KB), after which every
rank performs the identical reduction -- the result does not depend on the world
size.  Works with NCCL (GPU) and gloo (CPU tests).
"""

from __future__ import annotations


def shard_range(num_probes: int, world: int, rank: int):
    """Contiguous slice ``[p0, p1)`` of the probe range owned by `rank`."""
    per = -(-num_probes // world)
    return min(num_probes, rank * per), min(num_probes, (rank + 1) * per)


def gather_shards(local, num_probes: int, group=None):
    """All-gather the per-probe values of all ranks, in probe order."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    per = -(-num_probes // world)
    buf = torch.zeros((per,), dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    chunks = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(chunks, buf, group=group)
    parts = []
    for r in range(world):
        p0, p1 = shard_range(num_probes, world, r)
        parts.append(chunks[r][: p1 - p0])
    return torch.cat(parts)

"""Scene sharding across ranks and the final metadata gather (one process per GPU; scenes are
independent — /root/reference/gapro/gen_ps.py:36 is a plain sequential loop — so there is no
data-path collective, only one gather of per-scene records at the end)."""
from __future__ import annotations

from typing import List, Sequence


def shard_scenes(items: Sequence, rank: int, world: int, costs: Sequence[float] | None = None) -> List:
    """Items of this rank.  Without costs: round-robin over the sorted list (every rank sees the
    same list).  With costs: greedy longest-processing-time assignment, deterministic on all ranks."""
    if world <= 1:
        return list(items)
    if costs is None:
        return list(items[rank::world])
    order = sorted(range(len(items)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(costs[i])
        if r == rank:
            mine.append(i)
    return [items[i] for i in sorted(mine)]


def gather_records(local_records: list, world: int) -> list:
    """All ranks' per-scene records on every rank (torch.distributed all_gather_object: NCCL on the
    GPUs, gloo in the CPU tests)."""
    if world <= 1:
        return list(local_records)
    import torch.distributed as dist
    bucket = [None] * world
    dist.all_gather_object(bucket, list(local_records))
    out = []
    for part in bucket:
        out.extend(part)
    return out

"""Scene sharding across ranks and the final metadata gather (one process per GPU; scenes are
independent — /root/reference/gapro/gen_ps.py:36 is a plain sequential loop — so there is no
data-path collective, only small gathers of per-scene records: the cost estimates before the
labelling pass when cost balancing is on, the label metadata after it)."""
from __future__ import annotations

from typing import List, Sequence

# seconds of one B200 per unit of sum(M^3), sum(M^2), per point and per region, for turning the stage-pass statistics
# into a cost.  From the event-timed phases of the bench pass (sum M^3 = 9.85e10, sum M^2 = 9.5e7, 50 steps): tile
# products + Cholesky sweep 1445 ms ~ M^3, kernel build / gradient / column statistics 100 ms ~ M^2; the rest is per
# point / per launch.
COST_PER_M3 = 1.47e-11
COST_PER_M2 = 1.05e-9
COST_PER_POINT = 2e-8
COST_PER_REGION = 2e-5


def scene_cost(sum_m3: float, n_points: float = 0.0, n_regions: float = 0.0, sum_m2: float = 0.0) -> float:
    return (COST_PER_M3 * float(sum_m3) + COST_PER_M2 * float(sum_m2) + COST_PER_POINT * float(n_points) +
            COST_PER_REGION * float(n_regions))


def lpt_assignment(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment, deterministic on every rank: item indices per rank, each
    rank's list in DECREASING cost (heavy scenes first, so the tail of the job is made of light ones)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(costs[i])
        out[r].append(i)
    return out


def shard_scenes(items: Sequence, rank: int, world: int, costs: Sequence[float] | None = None) -> List:
    """Items of this rank.  Without costs: round-robin over the sorted list (every rank sees the
    same list).  With costs: LPT assignment (see lpt_assignment), items returned in list order."""
    if world <= 1:
        return list(items)
    if costs is None:
        return list(items[rank::world])
    return [items[i] for i in sorted(lpt_assignment(costs, world)[rank])]


def balance_stats(costs: Sequence[float], assignment: List[List[int]]) -> dict:
    loads = [sum(float(costs[i]) for i in part) for part in assignment]
    mean = sum(loads) / max(len(loads), 1)
    return dict(loads=loads, max_over_mean=(max(loads) / mean) if mean > 0 else 1.0)


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

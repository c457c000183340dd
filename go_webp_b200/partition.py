"""Batch -> GPU partitioning for the pixel hot path.

Images are independent, so sharding is static and needs no collective: equal
images are split into contiguous blocks, mixed sizes use greedy LPT (largest
processing time first) on the algorithmic byte cost.  Pure host logic (tested
with gloo world_size 2 on CPU).
"""
from __future__ import annotations

import heapq


def contiguous_shard(n_items: int, world_size: int, rank: int) -> range:
    """Items [lo, hi) of rank: block partition with the remainder spread over
    the first ranks."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return range(lo, hi)


def lpt_partition(costs, world_size: int):
    """Greedy LPT: returns a list (per rank) of item indices.  Deterministic:
    ties broken by index so every rank computes the same assignment."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    for lst in out:
        lst.sort()
    return out


def import_cost_bytes(w: int, h: int, with_alpha: bool = True) -> int:
    """Algorithmic bytes of one import: 4wh in, wh Y, 2*ceil(w/2)ceil(h/2) UV, wh A."""
    return 4 * w * h + w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + (w * h if with_alpha else 0)


def upsample_cost_bytes(w: int, h: int, with_alpha: bool = False, bpp: int = 4) -> int:
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) + bpp * w * h + (w * h if with_alpha else 0)


def size_class(w: int, h: int) -> str:
    """Class of an image for the calibrated cost model of mixed-size batches: the kernels' rate depends
    on the row alignment (16-byte rows take the vector paths, odd widths the general strips) and on how
    many full strips a row holds (narrow rows are mostly edge handling)."""
    al = "w16" if w % 16 == 0 else ("w4" if w % 4 == 0 else ("w2" if w % 2 == 0 else "w1"))
    sz = "narrow" if w < 512 else ("mid" if w < 2048 else "wide")
    return al + "/" + sz


def full_pipeline_cost_bytes(w: int, h: int) -> int:
    """import + fancy upsample with alpha plane + 8 B/px ARGB chain (configs[4])."""
    return import_cost_bytes(w, h) + upsample_cost_bytes(w, h, True) + 8 * w * h


def calibrated_costs(sizes, rates: dict, cost_bytes=full_pipeline_cost_bytes):
    """Per-image cost in TIME: algorithmic bytes / measured rate of the image's class (rates: class ->
    bytes per millisecond from one warm-up pass).  Classes without a measurement (or a zero one) fall
    back to the mean rate, so the model degrades to the byte model, never to nonsense."""
    good = [r for r in rates.values() if r > 0]
    mean = sum(good) / len(good) if good else 1.0
    return [cost_bytes(w, h) / (rates.get(size_class(w, h), 0.0) or mean) for (w, h) in sizes]

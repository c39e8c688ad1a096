"""Size-balanced region -> GPU assignment.

The reference parallelises by handing equal COUNTS of containers to independent Snakemake jobs
(crs_to_batches.build_batches, candidateregions/crs_to_batches.py:61-146; main.smk:634-641); DP work
per region varies by orders of magnitude with read length and coverage, so below that API regions are
assigned to GPUs by estimated cell count: longest-processing-time-first onto a min-heap of GPU loads.
Regions are independent: there is no collective on the data path, only a host-side gather of results.
"""
from __future__ import annotations

import heapq
from typing import Sequence

import numpy as np


def band_cells(m: int, n: int, band_lo: int, band_w: int) -> int:
    """Number of DP cells of the spec for one pair: |{(i,j): 1<=i<=m, 1<=j<=n, band_lo <= j-i < band_lo+band_w}|."""
    def below(x: int) -> int:            # cells with j - i <= x
        total = 0
        i0, i1 = max(1, 1 - x), min(m, n - x - 1)
        if i1 >= i0:
            total += (i1 - i0 + 1) * (i0 + i1) // 2 + x * (i1 - i0 + 1)
        i2 = max(1, n - x)
        if i2 <= m:
            total += n * (m - i2 + 1)
        return total
    return below(band_lo + band_w - 1) - below(band_lo - 1)


def region_cost(read_lengths: Sequence[int], slack: int = 300, granule: int = 128) -> int:
    """Estimated cells of the all-vs-all of one region from read lengths only (band = |dlen| + 2*slack,
    rounded up to the granule, times the shorter read)."""
    ln = np.sort(np.asarray(read_lengths, dtype=np.int64))
    n = len(ln)
    if n < 2:
        return 0
    a, b = np.triu_indices(n, k=1)
    short, diff = ln[a], ln[b] - ln[a]
    band = (diff + 2 * slack + granule) // granule * granule
    return int(np.sum(np.minimum(band, short + ln[b]) * short))


def lpt_assign(costs: Sequence[int], n_bins: int) -> list[list[int]]:
    """Longest-processing-time-first: item indices per bin, heaviest first. Deterministic:
    ties by lower index, then lower bin number."""
    order = sorted(range(len(costs)), key=lambda k: (-int(costs[k]), k))
    heap = [(0, b) for b in range(n_bins)]
    bins: list[list[int]] = [[] for _ in range(n_bins)]
    for k in order:
        load, b = heapq.heappop(heap)
        bins[b].append(k)
        heapq.heappush(heap, (load + int(costs[k]), b))
    return bins


def imbalance(costs: Sequence[int], bins: list[list[int]]) -> float:
    """max bin load / mean bin load (1.0 = perfect)."""
    loads = [sum(int(costs[k]) for k in b) for b in bins]
    mean = sum(loads) / max(len(loads), 1)
    return max(loads) / mean if mean > 0 else 1.0

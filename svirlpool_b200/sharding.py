"""Region sharding across the GPUs of one box: one process per GPU (torch.distributed for the plumbing),
regions assigned by the size-balanced scheduler, NO collective on the data path — each rank aligns its
regions independently and rank 0 gathers the per-region results on the host (the reference's
equivalent is one Snakemake job per batch whose JSONL outputs are concatenated, main.smk:634-641, :711-727).
"""
from __future__ import annotations

from typing import Callable, Sequence

from . import scheduler


def my_regions(costs: Sequence[int], rank: int, world: int) -> list[int]:
    """Indices of the regions this rank processes (same deterministic LPT assignment on every rank)."""
    return scheduler.lpt_assign(costs, world)[rank]


def run_sharded(region_ids: Sequence[int], costs: Sequence[int], process_region: Callable[[int], object] | None = None, dist=None,
                process_many: Callable[[list[int]], list[object]] | None = None):
    """Process `region_ids` across the ranks of the default process group and return, on rank 0, the list of
    per-region results in the order of region_ids (None on the other ranks). `dist` is torch.distributed
    (nccl on GPUs, gloo in the CPU tests); without it everything runs in this process. Either `process_region` (one region
    per call) or `process_many` (all regions of this rank in ONE call — one engine call per rank, e.g.
    consensus.ava_similarity_summary_many; results in the order of the ids it is given)."""
    if (process_region is None) == (process_many is None):
        raise ValueError("pass exactly one of process_region / process_many")
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = my_regions(costs, rank, world)
    if process_many is not None:
        results = process_many([region_ids[k] for k in mine]) if mine else []
        if len(results) != len(mine):
            raise ValueError("process_many must return one result per region id")
        local = dict(zip(mine, results))
    else:
        local = {k: process_region(region_ids[k]) for k in mine}
    if world == 1:
        return [local[k] for k in range(len(region_ids))]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)          # host-side gather of results, not a data-path collective
    if rank != 0:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[k] for k in range(len(region_ids))]

"""Chunk-grid sharding for the multi-GPU build.

Units = chunks; with one interner per GPU they are independent (the reference shares one store only
for dedup: world/voxmodel.rs:31-32), so the build needs NO data-path collective.  Two layouts:

* ``weak``   every rank builds its own ``grid`` of chunks, offset along X by ``rank * grid[0]`` chunks
             (what bench.py reports: per-GPU work is fixed as N grows);
* ``slab``   one fixed grid is split into contiguous X-slabs, ``g = cx * G // grid_x`` (keeps
             neighbouring terrain on one GPU, which maximises per-GPU dedup).
"""
from __future__ import annotations


def weak_shard(rank: int, grid=(64, 8, 64)):
    """-> (grid, x_chunk_offset) of the world rank ``rank`` builds."""
    return tuple(grid), rank * grid[0]


def slab_bounds(rank: int, world: int, grid_x: int):
    """Contiguous X-slab [lo, hi) of chunk columns owned by ``rank``: cx belongs to cx * world // grid_x."""
    lo = -(-rank * grid_x // world)          # ceil(rank * grid_x / world)
    hi = -(-(rank + 1) * grid_x // world)
    return lo, min(hi, grid_x)


def owner_of(cx: int, world: int, grid_x: int) -> int:
    return cx * world // grid_x


def chunk_linear_index(cx: int, cy: int, cz: int, grid) -> int:
    """Same order as workloads.terrain_world: (cx * gy + cy) * gz + cz."""
    return (cx * grid[1] + cy) * grid[2] + cz


def aggregate_throughput(chunks_per_rank, ms_per_rank):
    """Whole-job chunks/s: all ranks' chunks over the slowest rank's time (max over ranks)."""
    return sum(chunks_per_rank) / (max(ms_per_rank) * 1e-3)

"""N > 1 host logic on CPU: two gloo ranks shard a chunk grid, build their shards with per-rank
interners (the oracle stands in for the GPU here) and the gathered result must equal the
single-process build — same roots-to-voxels, every chunk owned exactly once, no collective on the
data path (only the gathers of results that bench.py also does)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from voxelis_b200 import sharding, workloads as wl
    grid = (4, 2, 3)
    # weak layout: each rank its own grid, offset along X
    g, xoff = sharding.weak_shard(rank, grid)
    masks, values = wl.terrain_world(g, 4, "surface_and_below", wl.U8, x_chunk_offset=xoff, materials=3)
    it = o.VoxInterner(16 << 20)
    roots, changed = it.apply_batches_fresh(4, masks, values)
    dense = np.stack([it.root_to_vec(r, 4) for r in roots])
    # gather what bench.py gathers: per-rank chunk counts and timings (fake ms = rank + 1)
    counts = [None] * world
    dist.all_gather_object(counts, int(masks.shape[0]))
    ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, dense)
    uniq = [None] * world
    dist.all_gather_object(uniq, int(it.next_index) - 1)
    if rank == 0:
        ret["counts"] = counts
        ret["ms_max"] = float(ms.item())
        ret["dense"] = np.concatenate(gathered)
        ret["unique_per_rank"] = uniq
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_weak_sharding_matches_single_process():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    from oracle import oracle as o
    from voxelis_b200 import sharding, workloads as wl
    grid = (4, 2, 3)
    # single process: the double-width world in one interner
    masks, values = wl.terrain_world((grid[0] * world, grid[1], grid[2]), 4, "surface_and_below", wl.U8, materials=3)
    it = o.VoxInterner(16 << 20)
    roots, _ = it.apply_batches_fresh(4, masks, values)
    dense = np.stack([it.root_to_vec(r, 4) for r in roots])
    assert ret["counts"] == [grid[0] * grid[1] * grid[2]] * world
    assert np.array_equal(ret["dense"], dense)          # rank r's chunks are columns [r*gx, (r+1)*gx)
    assert ret["ms_max"] == 2.0
    assert sharding.aggregate_throughput(ret["counts"], [1.0, 2.0]) == sum(ret["counts"]) / 2e-3
    # per-GPU interners never hold fewer nodes in total than one shared interner (dedup is per rank)
    assert sum(ret["unique_per_rank"]) >= it.next_index - 1


def test_slab_partition_covers_every_column_once():
    from voxelis_b200 import sharding
    for world in (1, 2, 4, 8):
        for gx in (8, 13, 64):
            owners = [sharding.owner_of(cx, world, gx) for cx in range(gx)]
            for r in range(world):
                lo, hi = sharding.slab_bounds(r, world, gx)
                assert [cx for cx in range(gx) if owners[cx] == r] == list(range(lo, hi))
            assert sorted(set(owners)) == [r for r in range(world) if sharding.slab_bounds(r, world, gx)[0] < sharding.slab_bounds(r, world, gx)[1]]

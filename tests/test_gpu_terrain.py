"""Device batch generation (vx_terrain_heights_device / vx_terrain_batches_device) against the numpy
generators of voxelis_b200/workloads.py (height_field, terrain_world — the inputs every other test and the bench feed
to oracle and GPU alike): byte-identical Batch arrays (reference layout core/batch.rs:39-45,153-157; patterns
utils/shapes.rs:273-357), and the build of the generated slab equals the oracle's build of the numpy slab."""
import numpy as np
import pytest
import torch

import parity
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def generate(g, depth, grid, dtype, variant, materials, x_chunk_offset=0):
    gx, gy, gz = grid
    n = 1 << depth
    B = wl.blocks_per_chunk(depth)
    dev = torch.device("cuda", 0)
    h = torch.empty((gx * n, gz * n), dtype=torch.int32, device=dev)
    m = torch.empty((gx * gy * gz, B, 2), dtype=torch.uint8, device=dev)
    v = torch.empty((gx * gy * gz, B, 8), dtype=torch.uint8 if dtype == wl.U8 else torch.int32, device=dev)
    torch.cuda.synchronize()
    g.terrain_heights_device(gx * n, gz * n, h.data_ptr(), wl.SEED_BASE, gy * n, x_chunk_offset * n, 0)
    g.terrain_batches_device(depth, grid, h.data_ptr(), m.data_ptr(), v.data_ptr(), variant == "surface_only", materials)
    g.sync()
    return h, m, v


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("variant,materials", [("surface_only", 1), ("surface_and_below", 1), ("surface_and_below", 3)])
def test_generated_batches_equal_numpy_generator(gpu_api, variant, materials, dtype):
    for depth, grid, xoff in ((5, (5, 3, 4), 0), (5, (2, 8, 2), 61), (4, (3, 2, 5), 7), (6, (2, 2, 1), 3)):
        g = gpu_api.VoxInterner.with_memory_budget(8 << 20, dtype)
        h, m, v = generate(g, depth, grid, dtype, variant, materials, xoff)
        n = 1 << depth
        H = wl.height_field(grid[0] * n, grid[2] * n, wl.SEED_BASE, height=grid[1] * n, x0=xoff * n)
        assert np.array_equal(h.cpu().numpy(), H), (depth, grid)
        em, ev = wl.terrain_world(grid, depth, variant, dtype, x_chunk_offset=xoff, materials=materials)
        assert np.array_equal(m.cpu().numpy(), em), (depth, grid)
        assert np.array_equal(v.cpu().numpy(), ev), (depth, grid)


def test_generated_world_builds_like_the_numpy_world(gpu_api, oracle_api):
    depth, grid = 5, (6, 4, 6)
    g = gpu_api.VoxInterner.with_memory_budget(64 << 20, wl.U8)
    _, m, v = generate(g, depth, grid, wl.U8, "surface_and_below", 3)
    nchunks = m.shape[0]
    roots = torch.zeros(nchunks, dtype=torch.int64, device=m.device)
    changed = torch.zeros(nchunks, dtype=torch.uint8, device=m.device)
    torch.cuda.synchronize()
    g.apply_batches_device(depth, nchunks, m.data_ptr(), v.data_ptr(), roots.data_ptr(), changed.data_ptr())
    g.sync()
    em, ev = wl.terrain_world(grid, depth, "surface_and_below", wl.U8, materials=3)
    c = oracle_api.VoxInterner(64 << 20, wl.U8)
    flags, fills = parity.flags_from(nchunks)
    croots, cchanged = c.apply_batches_fresh(depth, em, ev, flags & 1, fills, (flags >> 1) & 1)
    parity.assert_parity(gpu_api, oracle_api, depth, g, roots.cpu().numpy().astype(np.uint64),
                         changed.cpu().numpy(), c, croots, cchanged)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth,k,cell", [(5, 255, 1), (5, 4, 1), (6, 255, 4), (4, 17, 2)])
def test_random_batches_equal_numpy_generator(gpu_api, depth, k, cell, dtype):
    """vx_random_batches_device is workloads.p_random byte for byte (masks incl. the clear bits of k = 4)."""
    import torch
    vx = gpu_api
    n, c0 = 3, 5
    B = wl.blocks_per_chunk(depth)
    want_m, want_v = wl.batch_from_function(depth, wl.p_random(k, cell), dtype, n, chunk_arg=[c0 + i for i in range(n)])
    g = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    dm = torch.empty((n, B, 2), dtype=torch.uint8, device="cuda")
    dv = torch.empty((n, B, 8), dtype=torch.uint8 if dtype == wl.U8 else torch.int32, device="cuda")
    torch.cuda.synchronize()
    g.random_batches_device(depth, n, dm.data_ptr(), dv.data_ptr(), k, cell, chunk0=c0)
    g.sync()
    assert np.array_equal(dm.cpu().numpy(), want_m)
    assert np.array_equal(dv.cpu().numpy(), want_v)

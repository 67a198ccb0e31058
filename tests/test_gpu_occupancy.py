"""vx_occupancy_masks (CUDA, through the C ABI) == the oracle's restatement of generate_occupancy_masks +
OccupancyDataBuilder::build (reference voxelis/src/utils/mesh.rs:263-285,418-596): global planes, global_active,
sorted materials with voxel counts, per-material planes — bit-exact."""
import numpy as np
import pytest

import parity
from test_oracle_occupancy import cell_offsets, chunk_set
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def check_builders(g, groots, c, croots, depth, lod, groups, offs, M):
    """groups[b] = chunk indices of builder b; offs[b] = their offsets."""
    idx = [i for grp in groups for i in grp]
    flat_offs = [o for per in offs for o in per]
    bo = [b for b, grp in enumerate(groups) for _ in grp]
    got = g.occupancy_masks(groots[idx], depth, flat_offs, bo, len(groups), lod=lod, max_materials=M)
    for b, grp in enumerate(groups):
        want = c.occupancy_masks(croots[grp], depth, offs[b], lod=lod)
        nm = len(want["material_ids"])
        where = (depth, lod, b)
        assert got["n_materials"][b] == nm, where
        assert np.array_equal(got["material_ids"][b, :nm], want["material_ids"]), where
        assert np.array_equal(got["material_counts"][b, :nm], want["material_counts"]), where
        assert np.array_equal(got["global"][b], want["global"]), where
        assert np.array_equal(got["active"][b], want["active"]), where
        assert np.array_equal(got["per_material"][b, :nm], want["per_material"]), where


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5, 6])
def test_occupancy_masks_match_oracle(gpu_api, oracle_api, depth, dtype):
    masks, values = chunk_set(depth, dtype)
    g, groots, _, c, croots, _ = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype, budget=256 << 20)
    n = len(groots)
    for lod in range(0, depth + 1, 2 if depth > 3 else 1):
        S = 1 << (depth - lod)
        per = min((64 // S) ** 3, 5)
        groups = [list(range(b0, min(b0 + per, n))) for b0 in range(0, n, per)]
        offs = [cell_offsets(S, len(grp), seed=grp[0] + lod) for grp in groups]
        check_builders(g, groots, c, croots, depth, lod, groups, offs, 256)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5, 6])
def test_occupancy_single_material_fused_path(gpu_api, oracle_api, depth, dtype):
    """Single-material builders with max_materials == 1 (32 KiB of shared memory per CTA), every level of detail with
    chunks of at least 8^3, negative i32 material.  (A fused variant — one CTA per builder filling all three planes from
    one walk — passed this test too but measured only 4 % faster here and 10 % slower on the 3-material world.)"""
    parts = [wl.terrain_world((2, 2, 2), depth, "surface_and_below", dtype), wl.terrain_world((2, 1, 2), depth, "surface_only", dtype),
             wl.named_workload("uniform", 1, depth, dtype), wl.named_workload("hollow", 1, depth, dtype),
             wl.named_workload("checkerboard", 1, depth, dtype), wl.named_workload("sparse", 1, depth, dtype),
             wl.named_workload("diagonal", 1, depth, dtype), wl.named_workload("uniform_half", 1, depth, dtype)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    if dtype == wl.I32:
        values = (values * -5).astype(np.int32)                          # one negative material
    g, groots, _, c, croots, _ = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype, budget=256 << 20)
    n = len(groots)
    for lod in range(0, depth - 2):                                      # chunks of at least 8^3: the shared-memory path
        S = 1 << (depth - lod)
        per = min((64 // S) ** 3, 6)
        groups = [list(range(b0, min(b0 + per, n))) for b0 in range(0, n, per)]
        offs = [cell_offsets(S, len(grp), seed=grp[0] + lod) for grp in groups]
        check_builders(g, groots, c, croots, depth, lod, groups, offs, 1)


def test_occupancy_terrain_world_eight_chunks_per_builder(gpu_api, oracle_api):
    """The mesher's own packing (mesh.rs:598-606): 2x2x2 chunks of 32^3 per 64^3 volume."""
    grid = (4, 4, 4)
    masks, values = wl.terrain_world(grid, 5, "surface_and_below", wl.U8, materials=3)
    g, groots, _, c, croots, _ = parity.build_both(gpu_api, oracle_api, 5, masks, values, wl.U8)
    gx, gy, gz = grid
    groups, offs = [], []
    for bx in range(gx // 2):
        for by in range(gy // 2):
            for bz in range(gz // 2):
                grp, o = [], []
                for dx in range(2):
                    for dy in range(2):
                        for dz in range(2):
                            grp.append(((2 * bx + dx) * gy + 2 * by + dy) * gz + 2 * bz + dz)
                            o.append((32 * dx, 32 * dy, 32 * dz))
                groups.append(grp)
                offs.append(o)
    check_builders(g, groots, c, croots, 5, 0, groups, offs, 4)


def test_occupancy_whole_volume_leaf_and_empty(gpu_api, oracle_api):
    m, v = wl.named_workload("uniform", 2, 6, wl.U8)
    m[1] = 0
    g, groots, _, c, croots, _ = parity.build_both(gpu_api, oracle_api, 6, m, v, wl.U8)
    check_builders(g, groots, c, croots, 6, 0, [[0], [1]], [[(0, 0, 0)], [(0, 0, 0)]], 2)
    got = g.occupancy_masks(groots[:1], 6, [(0, 0, 0)], max_materials=1)
    assert (got["global"] == np.uint64(2**64 - 1)).all() and (got["active"] == np.uint64(2**64 - 1)).all()


def test_occupancy_argument_errors(gpu_api, oracle_api):
    masks, values = wl.batch_from_function(5, wl.p_random(255), wl.U8, 1)
    g, groots, _, _, _, _ = parity.build_both(gpu_api, oracle_api, 5, masks, values, wl.U8)
    with pytest.raises(gpu_api.VoxelisError) as e:       # 255 materials, room for 4
        g.occupancy_masks(groots, 5, [(0, 0, 0)], max_materials=4)
    assert e.value.code == -5
    with pytest.raises(gpu_api.VoxelisError):            # not a multiple of the chunk side
        g.occupancy_masks(groots, 5, [(16, 0, 0)])
    with pytest.raises(gpu_api.VoxelisError):            # two chunks in one cell
        g.occupancy_masks(np.repeat(groots, 2), 5, [(0, 0, 0), (0, 0, 0)])
    with pytest.raises(gpu_api.VoxelisError) as e:       # 128^3 does not fit an occupancy volume
        g.occupancy_masks(groots, 7, [(0, 0, 0)])
    assert e.value.code == -4


def test_occupancy_device_outputs_equal_host_outputs(gpu_api, oracle_api):
    """All-device output pointers (what a renderer keeping the planes on the GPU passes) give the same bytes."""
    import ctypes as C
    import torch
    masks, values = wl.terrain_world((2, 4, 2), 5, "surface_and_below", wl.U8, materials=3)
    g, groots, _, _, _, _ = parity.build_both(gpu_api, oracle_api, 5, masks, values, wl.U8)
    n = len(groots)
    idx = np.arange(n)
    cx, cy, cz = idx // 8, (idx // 2) % 4, idx % 2
    bo = np.ascontiguousarray(cy // 2, np.uint32)
    offs = np.ascontiguousarray(np.stack([cx * 32, (cy % 2) * 32, cz * 32], 1), np.uint32)
    nb, M = 2, 3
    host = g.occupancy_masks(groots, 5, offs, bo, nb, max_materials=M)
    dev = torch.device("cuda", 0)
    outs = [torch.full(s, -1, dtype=t, device=dev) for s, t in (((nb, 3 * 4096), torch.int64), ((nb, 6), torch.int64),
            ((nb,), torch.int32), ((nb, M), torch.int64), ((nb, M), torch.int64), ((nb, M, 3 * 4096), torch.int64))]
    torch.cuda.synchronize()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    r = np.ascontiguousarray(groots, np.uint64)
    rc = gpu_api.lib().vx_occupancy_masks(g.h, 5, 0, n, p(r), p(offs), p(bo), nb, M,
                                         *[C.c_void_p(t.data_ptr()) for t in outs])
    assert rc == 0
    nm = outs[2].cpu().numpy()
    assert np.array_equal(nm, host["n_materials"])
    assert np.array_equal(outs[0].cpu().numpy().view(np.uint64), host["global"])
    assert np.array_equal(outs[1].cpu().numpy().view(np.uint64), host["active"])
    for b in range(nb):
        k = int(nm[b])
        assert np.array_equal(outs[3][b, :k].cpu().numpy().view(np.uint64), host["material_ids"][b, :k])
        assert np.array_equal(outs[4][b, :k].cpu().numpy().view(np.uint64), host["material_counts"][b, :k])
        assert np.array_equal(outs[5][b, :k].cpu().numpy().view(np.uint64), host["per_material"][b, :k])
    with pytest.raises(gpu_api.VoxelisError):                     # outputs split between host and device memory
        gpu_api.lib().vx_occupancy_masks.restype = C.c_int
        rc = gpu_api.lib().vx_occupancy_masks(g.h, 5, 0, n, p(r), p(offs), p(bo), nb, M, C.c_void_p(outs[0].data_ptr()),
                                             p(host["active"]), p(host["n_materials"]), p(host["material_ids"]),
                                             p(host["material_counts"]), p(host["per_material"]))
        if rc < 0:
            raise gpu_api.VoxelisError(rc, gpu_api.lib().vx_last_error().decode())


def _popcount64(t):
    """Per-element population count of an int64 tensor holding u64 words (SWAR; masks undo the arithmetic shifts)."""
    m1, m2, m4 = 0x5555555555555555, 0x3333333333333333, 0x0F0F0F0F0F0F0F0F
    t = t - ((t >> 1) & m1)
    t = (t & m2) + ((t >> 2) & m2)
    t = (t + (t >> 4)) & m4
    return ((t * 0x0101010101010101) >> 56) & 0xFF


def _or_reduce(t):
    while t.shape[-1] > 1:
        h = t.shape[-1] // 2
        t = t[..., :h] | t[..., h:]
    return t[..., 0]


def test_occupancy_full_world_properties(gpu_api):
    """BASELINE.json's full world (64x8x64 chunks of 32^3, surface-and-below, 3 materials), generated, built and
    unfolded on the device into 4 096 builders; checked through properties that do not need the oracle at this size:
    per-material planes are pairwise disjoint and OR to the global plane; the three planes of a material hold the same
    number of bits = its voxel count; counts per material id sum to the number of such voxels in the generated
    batches; global_active is the OR of the words whose bits run along that axis (mesh.rs:451-461)."""
    import ctypes as C
    import torch
    depth, grid, M = 5, (64, 8, 64), 3
    gx, gy, gz = grid
    N, B, n = 32, 4096, gx * gy * gz
    dev = torch.device("cuda", 0)
    g = gpu_api.VoxInterner.with_memory_budget(256 << 20, wl.U8)
    h = torch.empty((gx * N, gz * N), dtype=torch.int32, device=dev)
    m = torch.empty((n, B, 2), dtype=torch.uint8, device=dev)
    v = torch.empty((n, B, 8), dtype=torch.uint8, device=dev)
    roots = torch.zeros(n, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    g.terrain_heights_device(gx * N, gz * N, h.data_ptr(), wl.SEED_BASE, gy * N)
    g.terrain_batches_device(depth, grid, h.data_ptr(), m.data_ptr(), v.data_ptr(), False, 3)
    g.apply_batches_device(depth, n, m.data_ptr(), v.data_ptr(), roots.data_ptr())
    g.sync()
    hroots = roots.cpu().numpy().astype(np.uint64)
    idx = np.arange(n)
    cx, cy, cz = idx // (gy * gz), (idx // gz) % gy, idx % gz
    bo = np.ascontiguousarray(((cx // 2) * (gy // 2) + cy // 2) * (gz // 2) + cz // 2, np.uint32)
    offs = np.ascontiguousarray(np.stack([(cx % 2) * N, (cy % 2) * N, (cz % 2) * N], 1), np.uint32)
    nb = n // 8
    glob = torch.empty((nb, 3, 4096), dtype=torch.int64, device=dev)
    active = torch.empty((nb, 6), dtype=torch.int64, device=dev)
    nmat = torch.empty(nb, dtype=torch.int32, device=dev)
    ids = torch.empty((nb, M), dtype=torch.int64, device=dev)
    counts = torch.empty((nb, M), dtype=torch.int64, device=dev)
    pm = torch.empty((nb, M, 3, 4096), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = gpu_api.lib().vx_occupancy_masks(g.h, depth, 0, n, p(hroots), p(offs), p(bo), nb, M,
                                         *[C.c_void_p(t.data_ptr()) for t in (glob, active, nmat, ids, counts, pm)])
    assert rc == 0, gpu_api.lib().vx_last_error()
    valid = torch.arange(M, device=dev)[None, :] < nmat[:, None]                     # [nb][M]
    pmv = torch.where(valid[:, :, None, None], pm, torch.zeros_like(pm))
    cv = torch.where(valid, counts, torch.zeros_like(counts))
    assert torch.equal(pmv[:, 0] | pmv[:, 1] | pmv[:, 2], glob)                      # OR of the materials = global
    assert not ((pmv[:, 0] & pmv[:, 1]) | (pmv[:, 0] & pmv[:, 2]) | (pmv[:, 1] & pmv[:, 2])).any()   # disjoint
    bits = _popcount64(pmv).sum(-1)                                                  # [nb][M][3]
    assert torch.equal(bits, cv[:, :, None].expand(-1, -1, 3))                       # every plane: the voxel count
    assert torch.equal(_popcount64(glob).sum(-1), cv.sum(1)[:, None].expand(-1, 3))
    for k in (1, 2, 3):                                                              # ids ascending, totals exact
        assert int(cv[valid & (ids == k)].sum()) == int((v == k).sum())
    assert bool(((ids[:, 1:] > ids[:, :-1]) | ~valid[:, 1:]).all())
    xm, ym, zm = _or_reduce(glob[:, 0]), _or_reduce(glob[:, 1]), _or_reduce(glob[:, 2])
    assert torch.equal(active, torch.stack([ym, zm, zm, xm, ym, xm], 1))
    assert int(nmat.max()) == 3 and int((nmat == 0).sum()) > 0                       # surface, deep and empty builders

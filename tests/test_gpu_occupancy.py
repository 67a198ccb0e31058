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

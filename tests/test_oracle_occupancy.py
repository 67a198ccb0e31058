"""The oracle's restatement of generate_occupancy_masks / fill_masks_for_region (reference
voxelis/src/utils/mesh.rs:418-596) against a plain-numpy statement of the same planes computed from dense volumes
(tests/occupancy_ref.py).  The reference holds no fixture for these planes (its mesh tests compare vertex counts of
the whole mesher), so this cross-check is what pins the restatement."""
import numpy as np
import pytest

import occupancy_ref as oref
from test_oracle_canonical import oracle_build
from voxelis_b200 import workloads as wl


def chunk_set(depth, dtype):
    parts = [wl.terrain_world((2, 2, 2), depth, "surface_and_below", dtype, materials=3),
             wl.batch_from_function(depth, wl.p_random(255), dtype, 1),
             wl.batch_from_function(depth, wl.p_random(4), dtype, 2),
             wl.named_workload("uniform", 1, depth, dtype), wl.named_workload("hollow", 1, depth, dtype),
             wl.named_workload("sum", 1, depth, dtype), wl.named_workload("checkerboard", 1, depth, dtype),
             wl.batch_from_function(depth, wl.p_random(255, cell=4), dtype, 1)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    if dtype == wl.I32:
        values = np.where(values == 2, -7, values).astype(np.int32)      # a negative value: usize sign extension
    return masks, values


def cell_offsets(S, count, seed):
    G = 64 // S
    cells = np.random.default_rng(seed).permutation(G ** 3)[:count]
    return [((c % G) * S, (c // G % G) * S, (c // (G * G)) * S) for c in cells]


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5, 6])
def test_oracle_occupancy_matches_dense_definition(oracle_api, depth, dtype):
    masks, values = chunk_set(depth, dtype)
    it, roots, _ = oracle_build(oracle_api, depth, masks, values, dtype, budget=256 << 20)
    n = len(roots)
    for lod in range(0, depth + 1, 2 if depth > 3 else 1):
        S = 1 << (depth - lod)
        per_builder = min((64 // S) ** 3, 5)
        for b0 in range(0, n, per_builder):
            idx = list(range(b0, min(b0 + per_builder, n)))
            offs = cell_offsets(S, len(idx), seed=b0 + lod)
            got = it.occupancy_masks(roots[idx], depth, offs, lod=lod)
            dense = [it.root_to_vec(int(roots[i]), depth, lod) for i in idx]
            oref.assert_same(got, oref.occupancy_from_dense(oref.place(dense, offs)), (depth, lod, b0))


def test_oracle_occupancy_whole_volume_leaf(oracle_api):
    """side == MAX_VOXELS_PER_AXIS takes the fill(u64::MAX) branch (mesh.rs:487-512)."""
    m, v = wl.named_workload("uniform", 1, 6, wl.U8)
    it, roots, _ = oracle_build(oracle_api, 6, m, v, wl.U8)
    got = it.occupancy_masks(roots, 6, [(0, 0, 0)])
    assert (got["global"] == np.uint64(2**64 - 1)).all() and (got["active"] == np.uint64(2**64 - 1)).all()
    assert got["material_ids"].tolist() == [1] and got["material_counts"].tolist() == [64 ** 3]
    assert (got["per_material"] == np.uint64(2**64 - 1)).all()


def test_oracle_occupancy_empty_root(oracle_api):
    it = oracle_api.VoxInterner(1 << 20, wl.U8)
    got = it.occupancy_masks(np.zeros(1, np.uint64), 5, [(0, 0, 0)])
    assert not got["global"].any() and not got["active"].any() and len(got["material_ids"]) == 0

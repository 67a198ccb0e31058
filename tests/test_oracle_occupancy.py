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


def random_box_volume(rng, n, materials):
    """A chunk of side n built from random axis-aligned boxes of random materials (and some carved holes): a mix of
    large uniform regions, thin plates and single voxels, i.e. nodes at every depth of the tree."""
    vol = np.zeros((n, n, n), np.int64)                                   # [x][y][z]
    for _ in range(int(rng.integers(1, 10))):
        lo = rng.integers(0, n, 3)
        hi = np.minimum(lo + rng.integers(1, n + 1, 3), n)
        vol[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 0 if rng.random() < 0.2 else int(rng.choice(materials))
    return vol


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_occupancy_random_boxes_oracle_numpy_and_kernel_index_code(oracle_api, dtype):
    """Seeded random box worlds at D = 4 and 5: the oracle's planes equal the dense numpy definition, and the kernels'
    host-steppable index code (both families, tests/cpp/occ_host_check.cu) equals the oracle."""
    from test_occupancy_host_step import _p, load_stepper
    step = load_stepper()
    rng = np.random.default_rng(20261017 + dtype)
    mats = [1, 2, 3, 200] if dtype == wl.U8 else [1, -3, 70000, 5]
    for depth in (4, 5):
        n = 1 << depth
        G = 64 >> depth
        vols = [random_box_volume(rng, n, mats) for _ in range(6)]
        batches = [wl.batch_from_dense(v.astype(wl.NP_DTYPE[dtype]), None, dtype) for v in vols]
        masks = np.stack([b[0] for b in batches])
        values = np.stack([b[1] for b in batches])
        it, roots, _ = oracle_build(oracle_api, depth, masks, values, dtype, budget=128 << 20)
        offs = cell_offsets(n, len(roots), seed=depth)
        got = it.occupancy_masks(roots, depth, offs)
        dense = [it.root_to_vec(int(r), depth) for r in roots]
        for d, v in zip(dense, vols):
            assert np.array_equal(d, np.transpose(v, (1, 2, 0)).astype(d.dtype))     # [y][z][x] of the [x][y][z] input
        oref.assert_same(got, oref.occupancy_from_dense(oref.place(dense, offs)), depth)
        dl = it.download()
        children = np.ascontiguousarray(dl["children"], np.uint64)
        vals = np.ascontiguousarray(dl["values"].astype(wl.NP_DTYPE[dtype]))
        cell = np.zeros(G ** 3, np.uint64)
        for r, (ox, oy, oz) in zip(roots, offs):
            cell[((oy >> depth) * G + (oz >> depth)) * G + (ox >> depth)] = r
        for entry in ("occ_host_check", "occ_host_check_planes"):
            M = 16
            ids, counts = np.zeros(M, np.uint64), np.zeros(M, np.uint64)
            glob = np.zeros(3 * 4096, np.uint64)
            pm = np.full((M, 3 * 4096), 0xDEADBEEF, np.uint64)
            nm = getattr(step, entry)(_p(children), _p(vals), dtype, _p(cell), depth, M, _p(ids), _p(counts), _p(glob), _p(pm))
            assert nm == len(got["material_ids"]), (entry, depth)
            assert np.array_equal(ids[:nm], got["material_ids"]) and np.array_equal(counts[:nm], got["material_counts"])
            assert np.array_equal(glob, got["global"]) and np.array_equal(pm[:nm], got["per_material"]), (entry, depth)

"""CUDA path vs oracle: voxel-exact, DAG-isomorphic, same counters."""
import numpy as np
import pytest

import parity
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu

PATTERNS = {
    "uniform": wl.p_uniform(1), "uniform_half": wl.p_uniform_half(1),
    "checkerboard_bench": wl.p_checkerboard_bench(1), "checkerboard_test": wl.p_checkerboard_test(),
    "sum": wl.p_sum(1), "sparse": wl.p_sparse(), "hollow": wl.p_hollow_cube(), "diagonal": wl.p_diagonal(),
    "gradient": wl.p_gradient(), "random255": wl.p_random(255), "random4": wl.p_random(4),
    "cell4": wl.p_random(255, cell=4),
}


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("name", sorted(PATTERNS))
def test_single_chunk_patterns(gpu_api, oracle_api, name, depth, dtype):
    masks, values = wl.batch_from_function(depth, PATTERNS[name], dtype, 1)
    r = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype)
    parity.assert_parity(gpu_api, oracle_api, depth, *r)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 4, 5])
def test_many_chunks_one_interner(gpu_api, oracle_api, depth, dtype):
    """apply_batches == serial application in index order, ids up to permutation."""
    parts = [wl.batch_from_function(depth, wl.p_random(4), dtype, 24),
             wl.batch_from_function(depth, wl.p_sum_per_chunk(), dtype, 16),
             wl.batch_from_function(depth, wl.p_random(255), dtype, 8),
             wl.named_workload("checkerboard", 16, depth, dtype),
             wl.named_workload("uniform", 5, depth, dtype)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    masks = np.concatenate([masks, masks[:7]])      # exact repeats -> same roots
    values = np.concatenate([values, values[:7]])
    masks[3] = 0                                      # an empty batch: changed == false
    values[3] = 0
    r = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype, budget=256 << 20)
    g, groots, gchanged = r[0], r[1], r[2]
    assert gchanged[3] == 0 and groots[3] == 0
    n0 = masks.shape[0] - 7
    same = np.array([0, 1, 2, 4, 5, 6])                # chunk 3 was blanked after the copy was taken
    assert np.array_equal(groots[same], groots[n0 + same])
    parity.assert_parity(gpu_api, oracle_api, depth, *r)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_terrain_world_small(gpu_api, oracle_api, dtype):
    for variant, mats in (("surface_only", 1), ("surface_and_below", 3)):
        masks, values = wl.terrain_world((6, 3, 6), 5, variant, dtype, materials=mats)
        r = parity.build_both(gpu_api, oracle_api, 5, masks, values, dtype, budget=256 << 20)
        parity.assert_parity(gpu_api, oracle_api, 5, *r)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5])
def test_fill_variants(gpu_api, oracle_api, depth, dtype):
    rng = np.random.default_rng(11)
    B = wl.blocks_per_chunk(depth)
    n = 6
    masks = np.zeros((n, B, 2), np.uint8)
    values = np.zeros((n, B, 8), wl.NP_DTYPE[dtype])
    sel = rng.random((n, B, 8)) < 0.04
    vals = rng.integers(1, 4, (n, B, 8))
    vals[vals == 3] = 7        # never equal to the fill value (the reference mis-counts there)
    values[sel] = vals[sel].astype(values.dtype)
    masks[:, :, 0] = (sel * (1 << np.arange(8))).sum(-1).astype(np.uint8)
    # chunk 4: fully set uniform blocks everywhere with a value != fill -> collapses to a leaf root
    masks[4, :, 0] = 0xFF
    values[4] = 9
    r = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype, fill=3)
    parity.assert_parity(gpu_api, oracle_api, depth, *r, indeg=False)
    # fill only (no patches): root is the fill leaf, ref 1
    r = parity.build_both(gpu_api, oracle_api, depth, masks[:2] * 0, values[:2] * 0, dtype, fill=5, has_patches=False)
    assert gpu_api.id_is_leaf(r[1][0]) and r[1][0] == r[1][1]
    parity.assert_parity(gpu_api, oracle_api, depth, *r)
    # fill + patches that all equal the fill -> changed == false, root untouched
    m2, v2 = masks[:1] * 0, values[:1] * 0
    m2[0, 5, 0] = 0b101
    v2[0, 5, 0] = v2[0, 5, 2] = 3
    r = parity.build_both(gpu_api, oracle_api, depth, m2, v2, dtype, fill=3)
    assert r[2][0] == 0 and r[1][0] == 0
    parity.assert_parity(gpu_api, oracle_api, depth, *r, indeg=False)


def test_known_answers_d5(gpu_api):
    """SURVEY §8(c) table, through the C ABI."""
    known = {"uniform": (0, 1, 4681, 4095, 1), "checkerboard_bench": (5, 1, 0, 21059, 6),
             "sum": (83, 94, 0, 37272, 177), "hollow": (87, 1, 0, 7393, 88), "diagonal": (5, 1, 0, 57, 6)}
    for name, (br, lf, col, hits, misses) in known.items():
        masks, values = wl.batch_from_function(5, PATTERNS[name], wl.U8, 1)
        g = gpu_api.VoxInterner.with_memory_budget(256 << 20)     # BASELINE.json config 1 budget
        roots, changed = g.apply_batches_slab(5, masks, values)
        st = g.stats()
        assert (st["branch_nodes"] - 1, st["leaf_nodes"], st["collapsed_branches"]) == (br, lf, col)
        assert (st["total_cache_hits"], st["total_cache_misses"]) == (hits, misses)
        assert changed[0] == 1
        assert g.capacity == 3397917


def test_host_and_device_inputs_agree(gpu_api):
    import torch
    masks, values = wl.batch_from_function(5, wl.p_random(4), wl.U8, 64)
    a = gpu_api.VoxInterner.with_memory_budget(64 << 20)
    ra, ca = a.apply_batches_slab(5, masks, values)
    b = gpu_api.VoxInterner.with_memory_budget(64 << 20)
    dm, dv = torch.from_numpy(masks).cuda(), torch.from_numpy(values).cuda()
    droots = torch.zeros(64, dtype=torch.int64, device="cuda")
    dch = torch.zeros(64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    b.apply_batches_device(5, 64, dm.data_ptr(), dv.data_ptr(), droots.data_ptr(), dch.data_ptr())
    b.sync()
    assert np.array_equal(a.roots_to_vec(ra, 5), b.roots_to_vec(droots.cpu().numpy().astype(np.uint64), 5))
    assert np.array_equal(ca, dch.cpu().numpy())
    assert a.stats() == b.stats()


def test_out_of_memory_is_a_status_not_an_abort(gpu_api):
    g = gpu_api.VoxInterner.with_memory_budget(79 * 64)          # 64 nodes
    masks, values = wl.batch_from_function(5, wl.p_random(255), wl.U8, 1)
    with pytest.raises(gpu_api.VoxelisError) as e:
        g.apply_batches_slab(5, masks, values)
    assert e.value.code == -2 and "Out of memory" in str(e.value)   # interner/macros.rs:38
    with pytest.raises(gpu_api.VoxelisError) as e:
        g.apply_batches_slab(5, masks, values)
    assert e.value.code == -7
    g.reset()
    m2, v2 = wl.batch_from_function(5, wl.p_uniform(3), wl.U8, 1)
    roots, changed = g.apply_batches_slab(5, m2, v2)
    assert changed[0] == 1 and gpu_api.id_is_leaf(roots[0])


def test_bounds_and_argument_errors(gpu_api):
    vx = gpu_api
    it = vx.VoxInterner.with_memory_budget(1 << 20)
    t = vx.VoxTree(5)
    b = t.create_batch()
    with pytest.raises(vx.VoxelisError) as e:
        b.set(it, (32, 0, 0), 1)
    assert e.value.code == -5
    with pytest.raises(vx.VoxelisError) as e:
        t.get(it, (0, -1, 0))
    assert e.value.code == -5
    with pytest.raises(vx.VoxelisError):
        vx.VoxTree(8)
    with pytest.raises(vx.VoxelisError):
        vx.VoxInterner.with_memory_budget(10)
    assert t.get(it, (1, 2, 3)) is None and t.is_empty() and not t.is_dirty()


@pytest.mark.parametrize("run", ["1", "8"])
@pytest.mark.parametrize("depth", [5, 6])
def test_both_join_paths(gpu_api, oracle_api, depth, run, monkeypatch):
    """apply_kernel hands out runs of 8 units (a warp owns a 32^3 cube, local join) when there are many
    cubes, single units (atomic last-arriver join) otherwise; VX_FORCE_RUN pins either path."""
    monkeypatch.setenv("VX_FORCE_RUN", run)
    n = 24 if depth == 5 else 4
    parts = [wl.batch_from_function(depth, wl.p_random(4), wl.U8, n),
             wl.batch_from_function(depth, wl.p_random(255, cell=4), wl.U8, 3),
             wl.named_workload("hollow", 2, depth, wl.U8)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    masks[1] = 0
    r = parity.build_both(gpu_api, oracle_api, depth, masks, values, wl.U8, budget=512 << 20)
    parity.assert_parity(gpu_api, oracle_api, depth, *r)


def test_depth7_extension(gpu_api, oracle_api):
    """D = 7 (128^3) is beyond the reference (MaxDepth::new asserts < 7, core/max_depth.rs:77-83);
    parity is against the oracle generalised with the PATH_MASKS formula."""
    parts = [wl.batch_from_function(7, wl.p_random(255, cell=4), wl.U8, 1),
             wl.batch_from_function(7, wl.p_sparse(3, 8), wl.U8, 1)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    r = parity.build_both(gpu_api, oracle_api, 7, masks, values, wl.U8, budget=512 << 20)
    parity.assert_parity(gpu_api, oracle_api, 7, *r, dense_check=False)
    g, groots, c, croots = r[0], r[1], r[3], r[4]
    assert np.array_equal(g.roots_to_vec(groots[:1], 7)[0], c.root_to_vec(croots[0], 7))


def test_full_size_configs_properties(gpu_api, oracle_api):
    """BASELINE.json configs at their full sizes, checked through size-independent properties:
    config 2 — 4096 identical checkerboard / set_sum chunks in one interner: every root identical, node
    counts equal the single-chunk known answers (all-hit after the first chunk);
    config 3 — the 64x8x64 perlin-dunes world: changed flags == "batch has a set bit", unique node counts
    and hit/miss counters equal the oracle's, random chunks unfold voxel-exactly."""
    vx, o = gpu_api, oracle_api
    for name, (br, lf) in {"checkerboard": (5, 1), "sum": (83, 94)}.items():
        masks, values = wl.named_workload(name, 4096)
        g = vx.VoxInterner.with_memory_budget(256 << 20)
        roots, changed = g.apply_batches_slab(5, masks, values)
        assert changed.all() and (roots == roots[0]).all()
        st = g.stats()
        assert (st["branch_nodes"] - 1, st["leaf_nodes"]) == (br, lf)
        assert g.get_ref(int(roots[0])) == 4096                      # 4096 tree handles on one root
    masks, values = wl.terrain_world((64, 8, 64), 5, "surface_only")
    g = vx.VoxInterner.with_memory_budget(256 << 20)
    roots, changed = g.apply_batches_slab(5, masks, values)
    assert np.array_equal(changed.astype(bool), masks[:, :, 0].any(1))
    assert ((roots == 0) == (changed == 0)).all()
    c = o.VoxInterner(256 << 20)
    croots, cchanged = c.apply_batches_fresh(5, masks, values)
    gst, cst = g.stats(), c.stats()
    for k in ("branch_nodes", "leaf_nodes", "collapsed_branches", "total_cache_hits", "total_cache_misses"):
        assert gst[k] == cst[k], k
    rng = np.random.default_rng(0)
    pick = rng.choice(np.flatnonzero(changed), 24, replace=False)
    dense = g.roots_to_vec(roots[pick], 5)
    for j, i in enumerate(pick):
        assert np.array_equal(dense[j], wl.dense_expected(masks[i], values[i]))
    # identical batches anywhere in the world got identical roots (dedup across chunks)
    key = {}
    for i in np.flatnonzero(changed)[:2000]:
        k = (masks[i].tobytes(), values[i].tobytes())
        assert key.setdefault(k, roots[i]) == roots[i]


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_to_vec_at_every_lod(gpu_api, oracle_api, dtype):
    """vx_roots_to_vec_lod == to_vec(interner, root, max_depth.for_lod(lod)) (world/voxchunk.rs:267)."""
    for depth in (3, 5):
        parts = [wl.batch_from_function(depth, wl.p_random(4), dtype, 3),
                 wl.batch_from_function(depth, wl.p_random(255, cell=2), dtype, 2),
                 wl.named_workload("hollow", 1, depth, dtype), wl.named_workload("uniform", 1, depth, dtype),
                 wl.batch_from_function(depth, wl.p_sparse(), dtype, 2)]
        masks = np.concatenate([p[0] for p in parts])
        values = np.concatenate([p[1] for p in parts])
        masks[1] = 0
        g, groots, _, c, croots, _ = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype)
        for lod in range(depth + 2):
            gd = g.roots_to_vec(groots, depth, lod)
            for i in range(len(groots)):
                assert np.array_equal(gd[i], c.root_to_vec(int(croots[i]), depth, lod)), (depth, lod, i)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [5, 6])
def test_solid_units_take_the_short_way_with_the_same_counters(gpu_api, oracle_api, depth, dtype):
    """apply_kernel answers a unit whose 512 blocks are all set to one value with that value's leaf (solid_unit,
    vx_build.cuh) — DAG, refcounts and every counter must be those of the block-by-block build: chunks that are
    solid throughout, solid but for one voxel / one cleared bit / one unit, two solid halves, and the same under a fill."""
    B = wl.blocks_per_chunk(depth)
    n = 7
    masks = np.zeros((n, B, 2), np.uint8)
    values = np.zeros((n, B, 8), wl.NP_DTYPE[dtype])
    masks[:, :, 0] = 0xFF
    values[:] = 5
    values[1, 700 % B, 3] = 6                 # one voxel differs inside an otherwise solid unit
    masks[2, 513 % B, 0] = 0x7F               # one voxel not set
    values[2, 513 % B, 7] = 0
    values[3, :512] = 9                       # first unit solid with another value
    values[4, B // 2:] = 2                    # two solid halves
    values[5, 1024 % B: 1024 % B + 512] = 0   # one unit entirely unset
    masks[5, 1024 % B: 1024 % B + 512, 0] = 0
    values[6, 100 % B] = 7                    # one uniform block of another value inside a solid unit
    def check(label, *a, **kw):
        r = parity.build_both(gpu_api, oracle_api, depth, *a, **kw)
        try:
            parity.assert_parity(gpu_api, oracle_api, depth, *r, indeg=kw.get("fill") is None)
        except AssertionError as e:
            gd, cd = r[0].download(), r[3].download()
            raise AssertionError(f"{label}: gpu refs {sorted(gd['refs'][:gd['n']].tolist())[-6:]} values {gd['values'][:12].tolist()} "
                                 f"oracle refs {sorted(cd['refs'][:cd['n']].tolist())[-6:]} values {cd['values'][:12].tolist()}") from e
    for fill in (None, 3):   # never a fill equal to a set value: the reference's refcounts go wrong there
        check(f"fill={fill}", masks, values, dtype, fill=fill, budget=256 << 20)
    # one chunk at a time (n = 1 spreads the units of a chunk over warps that join through global memory)
    for i in (0, 1, 3):
        check(f"chunk {i} alone", masks[i:i + 1], values[i:i + 1], dtype)

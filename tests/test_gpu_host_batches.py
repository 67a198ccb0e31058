"""vx_apply_batches on HOST batch handles: only the touched units cross the bus (vx_stage.cuh).

Checked against the CPU oracle (same batches through the reference's serial loop,
voxelis-voxelize/src/lib.rs:357-361) and against the slab entry on the same arrays.  VX_STAGE_POISON
fills the device slab's values with garbage first, so a kernel that consumed the values of a block
without a set bit (the reference never does, spatial/voxtree.rs:779-781) would fail here."""
import os

import numpy as np
import pytest

import parity
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def poison():
    os.environ["VX_STAGE_POISON"] = "1"
    yield
    os.environ.pop("VX_STAGE_POISON", None)


def handles_from_arrays(vx, depth, dtype, masks, values, raw=False):
    """raw=False: Batch.assign (the batch stays API-only: the one-bit-per-block map travels);
    raw=True: the caller writes the arrays itself and calls mark_patched (the masks travel).
    Either way values under unset bits are junk the build must never look at."""
    rng = np.random.default_rng(5)
    trees = [vx.VoxTree(depth, dtype) for _ in range(masks.shape[0])]
    batches = [t.create_batch() for t in trees]
    for b, m, v in zip(batches, masks, values):
        bits = ((m[:, 0][:, None] >> np.arange(8)) & 1).astype(bool)
        junk = np.where(bits, v, rng.integers(1, 100, v.shape).astype(v.dtype))
        if raw:
            b.masks[:] = m
            b.values[:] = junk
            b.mark_patched()
        else:
            b.assign(m, junk)
    return vx.ChunkSet(trees, batches)


def world(depth, dtype):
    if depth == 5:
        parts = [wl.terrain_world((4, 2, 4), 5, "surface_only", dtype),
                 wl.terrain_world((2, 2, 2), 5, "surface_and_below", dtype, materials=3),
                 wl.batch_from_function(5, wl.p_random(255), dtype, 2),
                 wl.named_workload("checkerboard", 1, 5, dtype),
                 wl.named_workload("sum", 1, 5, dtype)]
    else:
        parts = [wl.batch_from_function(depth, wl.p_random(4), dtype, 3),
                 wl.named_workload("sum", 2, depth, dtype)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    # a few chunks with nothing in them, and one whose only voxels sit in the last unit
    masks[1] = 0
    values[1] = 0
    masks[2, :-8] = 0
    values[2, :-8] = 0
    return masks, values


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("builder", ["fused", "bulk"])
@pytest.mark.parametrize("raw", [False, True], ids=["api", "raw"])
def test_handles_match_oracle_and_slab(gpu_api, oracle_api, depth, dtype, builder, raw, monkeypatch):
    vx, o = gpu_api, oracle_api
    monkeypatch.setenv("VX_BUILDER", builder)
    masks, values = world(depth, dtype)
    n = masks.shape[0]
    cs = handles_from_arrays(vx, depth, dtype, masks, values, raw)
    g = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    gchanged = cs.apply(g).copy()
    groots = cs.roots()
    c = o.VoxInterner(64 << 20, dtype)
    flags, fills = parity.flags_from(n)
    # a batch nobody wrote to has has_patches == false and the reference answers "changed" for it with the
    # root still EMPTY (voxtree.rs:756-758, :303-328); assign() leaves an all-zero chunk in that state
    hp = np.array([b.has_patches for b in cs.batches], np.uint8)
    assert hp[1] == (1 if raw else 0) and hp.any()
    croots, cchanged = c.apply_batches_fresh(depth, masks, values, flags & 1, fills, hp)
    groots = np.where(gchanged.astype(bool), groots, croots)     # an unchanged tree keeps its (empty) root
    parity.assert_parity(vx, o, depth, g, groots, gchanged, c, croots, cchanged)
    # the slab entry on the same arrays: same counters, same voxels
    g2 = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    r2, ch2 = g2.apply_batches_slab(depth, masks, values)
    assert np.array_equal(ch2[hp == 1], gchanged[hp == 1])
    for k, v in g.stats().items():
        assert g2.stats()[k] == v, k
    assert np.array_equal(g.roots_to_vec(groots, depth), g2.roots_to_vec(np.where(ch2.astype(bool), r2, 0), depth))


def test_small_slices(gpu_api, oracle_api, monkeypatch):
    """VX_STAGE_MAX_BYTES forces several slices through one device slab."""
    vx, o = gpu_api, oracle_api
    monkeypatch.setenv("VX_STAGE_MAX_BYTES", str(2 * (5 * 40960 + 600)))
    masks, values = world(5, wl.U8)
    n = masks.shape[0]
    cs = handles_from_arrays(vx, 5, wl.U8, masks, values)
    g = vx.VoxInterner.with_memory_budget(64 << 20, wl.U8)
    gchanged = cs.apply(g).copy()
    c = o.VoxInterner(64 << 20, wl.U8)
    flags, fills = parity.flags_from(n)
    hp = np.array([b.has_patches for b in cs.batches], np.uint8)
    croots, cchanged = c.apply_batches_fresh(5, masks, values, flags & 1, fills, hp)
    groots = np.where(gchanged.astype(bool), cs.roots(), croots)
    parity.assert_parity(vx, o, 5, g, groots, gchanged, c, croots, cchanged)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_set_many_tracks_units(gpu_api, oracle_api, dtype):
    """Batch::set through the array entry: unit summary, clear, re-use, zero voxels (clear bits)."""
    vx, o = gpu_api, oracle_api
    rng = np.random.default_rng(7)
    depth, n = 5, 6
    g, c = vx.VoxInterner.with_memory_budget(32 << 20, dtype), o.VoxInterner(32 << 20, dtype)
    gtrees = [vx.VoxTree(depth, dtype) for _ in range(n)]
    gb = [t.create_batch() for t in gtrees]
    for rnd in range(3):
        otrees = [o.VoxTree(depth, dtype) for _ in range(n)]
        for i in range(n):
            k = int(rng.integers(0, 400)) if i else 0          # chunk 0 stays empty
            lo = rng.integers(0, 17, 3)
            xyz = (lo + rng.integers(0, 16, (k, 3))).astype(np.int32)   # inside a 16^3 corner: few units
            vals = rng.integers(0, 4, k).astype(np.int64)               # zeros record clear bits
            ob = otrees[i].create_batch()
            for p, v in zip(xyz, vals):
                ob.set(c, p, int(v))
            if k:
                gb[i].set_many(xyz, vals)
                assert 1 <= gb[i].touched_units <= 8
            assert gb[i].touched_units <= 8
            otrees[i].apply_batch(c, ob)
        if rnd:
            vx.trees_forget(gtrees)
        cs = vx.ChunkSet(gtrees, gb)
        cs.apply(g)
        for i in range(n):
            assert np.array_equal(gtrees[i].to_vec(g), otrees[i].to_vec(c)), (rnd, i)
        for b in gb:
            b.clear()
            assert b.touched_units == 0 and not b.has_patches and b.size() == 0
        g.reset()
        c = o.VoxInterner(32 << 20, dtype)


def test_fill_and_edit_through_handles(gpu_api, oracle_api):
    """fills and non-empty trees still go through the staged slab (flags / old roots)."""
    vx, o = gpu_api, oracle_api
    depth, dtype, n = 4, wl.U8, 5
    rng = np.random.default_rng(3)
    g, c = vx.VoxInterner.with_memory_budget(32 << 20, dtype), o.VoxInterner(32 << 20, dtype)
    gt = [vx.VoxTree(depth, dtype) for _ in range(n)]
    ot = [o.VoxTree(depth, dtype) for _ in range(n)]
    for rnd in range(4):
        gbs, obs = [], []
        for i in range(n):
            gb_, ob_ = gt[i].create_batch(), ot[i].create_batch()
            if i % 3 == 0:       # always the same trees: editing a filled tree without a new fill walks into a
                gb_.fill(g, 5)   # collapsed leaf, which the reference answers with a panic (block_id.rs types())
                ob_.fill(c, 5)
            for _ in range(int(rng.integers(0, 60))):
                p = rng.integers(0, 16, 3)
                v = int(rng.integers(1, 4))
                gb_.set(g, p, v)
                ob_.set(c, p, v)
            gbs.append(gb_)
            obs.append(ob_)
        gch = vx.apply_batches(g, gt, gbs)
        och = [ot[i].apply_batch(c, obs[i]) for i in range(n)]
        assert list(gch) == och
        for i in range(n):
            assert np.array_equal(gt[i].to_vec(g), ot[i].to_vec(c)), (rnd, i)


def _many_small(vx, o, depth, n, seed):
    """n small chunks filled through set_many, the same voxels recorded in oracle batches."""
    rng = np.random.default_rng(seed)
    N = 1 << depth
    gtrees = [vx.VoxTree(depth, wl.U8) for _ in range(n)]
    gb = [t.create_batch() for t in gtrees]
    ob = []
    oi = o.VoxInterner(64 << 20, wl.U8)
    for i in range(n):
        k = int(rng.integers(0, 6))
        xyz = rng.integers(0, N, (k, 3)).astype(np.int32)
        vals = rng.integers(1, 4, k).astype(np.int64)
        b = o.Batch(depth, wl.U8)
        for p, v in zip(xyz, vals):
            b.set(oi, p, int(v))
        if k:
            gb[i].set_many(xyz, vals)
        ob.append(b)
    return gtrees, gb, ob, oi


def test_big_call_is_sliced_and_late_errors_leave_the_interner_alone(gpu_api, oracle_api):
    """n >= 8192: the first slice's bus traffic is queued before the remaining handles are checked."""
    vx, o = gpu_api, oracle_api
    depth, n = 3, 9000
    gtrees, gb, ob, oi = _many_small(vx, o, depth, n, 11)
    g = vx.VoxInterner.with_memory_budget(64 << 20, wl.U8)
    # a handle of the wrong depth far into the call: VX_E_INVALID, nothing applied
    bad = vx.VoxTree(4, wl.U8).create_batch()
    with pytest.raises(vx.VoxelisError):
        vx.apply_batches(g, gtrees, gb[:8500] + [bad] + gb[8501:])
    assert g.stats()["alive_nodes"] == 1 and all(t.is_empty() for t in gtrees[::97])
    # the real call
    cs = vx.ChunkSet(gtrees, gb)
    changed = cs.apply(g).copy()
    otrees = [o.VoxTree(depth, wl.U8) for _ in range(n)]
    och = [otrees[i].apply_batch(oi, ob[i]) for i in range(n)]
    assert list(changed.astype(bool)) == och
    dense = g.roots_to_vec(cs.roots(), depth)
    for i in range(0, n, 7):
        assert np.array_equal(dense[i], otrees[i].to_vec(oi)), i
    assert g.stats()["alive_nodes"] == oi.stats()["alive_nodes"]


def test_repeated_tree_in_a_late_slice_takes_the_serial_loop(gpu_api, oracle_api):
    vx, o = gpu_api, oracle_api
    depth, n = 3, 8300
    gtrees, gb, ob, oi = _many_small(vx, o, depth, n, 12)
    g = vx.VoxInterner.with_memory_budget(64 << 20, wl.U8)
    gtrees[8250] = gtrees[10]                       # batch 10, then batch 8250, on the same tree
    vx.apply_batches(g, gtrees, gb)
    otrees = [o.VoxTree(depth, wl.U8) for _ in range(n)]
    otrees[8250] = otrees[10]
    for i in range(n):
        otrees[i].apply_batch(oi, ob[i])
    for i in (9, 10, 11, 8249, 8251, 8299):
        assert np.array_equal(gtrees[i].to_vec(g), otrees[i].to_vec(oi)), i


def test_untouched_batch_reports_changed_like_the_reference(gpu_api, oracle_api):
    """voxtree.rs:756-758 hands back the initial node when the batch has no patches; for an EMPTY tree that
    is EMPTY != INVALID, so apply_batch answers true and marks the tree dirty.  All three entries agree."""
    vx, o = gpu_api, oracle_api
    g, c = vx.VoxInterner.with_memory_budget(16 << 20), o.VoxInterner(16 << 20)
    gt, ot = vx.VoxTree(4), o.VoxTree(4)
    assert gt.apply_batch(g, gt.create_batch()) is True and ot.apply_batch(c, ot.create_batch()) is True
    assert gt.is_empty() and gt.is_dirty() and ot.is_empty() and ot.is_dirty()
    trees = [vx.VoxTree(4) for _ in range(3)]
    assert list(vx.apply_batches(g, trees, [t.create_batch() for t in trees])) == [True] * 3
    assert all(t.is_empty() and t.is_dirty() for t in trees)
    B = wl.blocks_per_chunk(4)
    roots, changed = g.apply_batches_slab(4, np.zeros((2, B, 2), np.uint8), np.zeros((2, B, 8), np.uint8),
                                          flags=np.zeros(2, np.uint8), fills=np.zeros(2, np.int64))
    assert list(changed) == [1, 1] and list(roots) == [0, 0]
    assert g.stats()["alive_nodes"] == 1


# ---------------------------------------------------------------------------------------------- journal path
def _set_world(vx, depth, dtype, n, seed, dense_every=0):
    """n batches written voxel by voxel through Batch::set (repeats, later clears, rewrites), returned with the dense
    volumes they must build: sparse ones keep a journal, every `dense_every`-th one overflows it."""
    rng = np.random.default_rng(seed)
    N = 1 << depth
    trees = [vx.VoxTree(depth, dtype) for _ in range(n)]
    batches = [t.create_batch() for t in trees]
    vols = np.zeros((n, N, N, N), np.int64)                                    # [i][x][y][z]
    for i, b in enumerate(batches):
        k = (N ** 3 // 3) if dense_every and i % dense_every == dense_every - 1 else int(rng.integers(0, 300))
        xyz = rng.integers(0, N, (k, 3))
        vals = rng.integers(-3 if dtype == wl.I32 else 0, 6, k)               # zeros are clears (batch.rs:165-168)
        for (x, y, z), v in zip(xyz, vals):
            b.set(None, (int(x), int(y), int(z)), int(v))
            vols[i, x, y, z] = v
        for (x, y, z) in xyz[: k // 4]:                                        # a second write to a quarter of them
            b.set(None, (int(x), int(y), int(z)), 4)
            vols[i, x, y, z] = 4
    return trees, batches, vols


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5])
def test_journal_path_matches_dense_expectation(gpu_api, depth, dtype, monkeypatch):
    """Batches written through set() travel as packed (block, values) journals; ones that outgrow the journal fall back
    to the occupancy bitmap for their whole slice.  Voxels must equal what was set, whichever path a slice took, and
    both staging paths must build the same DAG."""
    vx = gpu_api
    outs = {}
    for mode in ("journal", "occupancy", "mixed"):
        if mode == "occupancy":
            monkeypatch.setenv("VX_STAGE_NO_JOURNAL", "1")
        else:
            monkeypatch.delenv("VX_STAGE_NO_JOURNAL", raising=False)
        trees, batches, vols = _set_world(vx, depth, dtype, 40, seed=depth * 10 + dtype, dense_every=7 if mode == "mixed" else 0)
        g = vx.VoxInterner.with_memory_budget(128 << 20, dtype)
        changed = vx.apply_batches(g, trees, batches)
        for i, t in enumerate(trees):
            want = np.transpose(vols[i], (1, 2, 0))                            # [y][z][x]
            assert np.array_equal(t.to_vec(g).astype(np.int64), want), (mode, i)
            assert bool(changed[i]) == bool((vols[i] != 0).any()) or not (vols[i] != 0).any()
        outs[mode] = g.stats()["alive_nodes"]
        # edits on top (the same handles, cleared and refilled): the journal restarts with the batch
        for b in batches[:5]:
            b.clear()
            b.set(None, (1, 1, 1), 9)
        vx.apply_batches(g, trees[:5], batches[:5])
        for t, v in zip(trees[:5], vols[:5]):
            v2 = v.copy()
            v2[1, 1, 1] = 9
            assert np.array_equal(t.to_vec(g).astype(np.int64), np.transpose(v2, (1, 2, 0)))
    assert outs["journal"] == outs["occupancy"]

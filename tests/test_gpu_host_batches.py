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
    croots, cchanged = c.apply_batches_fresh(depth, masks, values, flags & 1, fills, (flags >> 1) & 1)
    groots = np.where(gchanged.astype(bool), groots, croots)     # an unchanged tree keeps its (empty) root
    parity.assert_parity(vx, o, depth, g, groots, gchanged, c, croots, cchanged)
    # the slab entry on the same arrays: same counters, same voxels
    g2 = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    r2, ch2 = g2.apply_batches_slab(depth, masks, values)
    assert np.array_equal(ch2, gchanged)
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
    croots, cchanged = c.apply_batches_fresh(5, masks, values, flags & 1, fills, (flags >> 1) & 1)
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

"""apply_batch on NON-EMPTY trees + release: old-tree merge (voxtree.rs:785-842,
:930-952), dec_ref_recursive / recycle / generations / free list (interner/mod.rs:419-625)."""
import numpy as np
import pytest

import parity

from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def set_batch_arrays(batch, masks, values):
    batch.masks[:] = masks
    batch.values[:] = values
    if hasattr(batch, "mark_patched"):
        batch.mark_patched()
    else:
        batch.has_patches = True


def random_edit(rng, depth, dtype, density, alphabet, uniform_blocks=0.1):
    B = wl.blocks_per_chunk(depth)
    sel = rng.random((B, 8)) < density
    vals = rng.choice(alphabet, (B, 8))
    full = rng.random(B) < uniform_blocks          # some blocks fully set to one value (collapse path)
    sel[full] = True
    vals[full] = rng.choice(alphabet, (int(full.sum()), 1))
    masks = np.zeros((B, 2), np.uint8)
    masks[:, 0] = (sel * (1 << np.arange(8))).sum(1).astype(np.uint8)
    values = np.where(sel, vals, 0).astype(wl.NP_DTYPE[dtype])
    return masks, values


def overlay(dense_yzx, masks, values, depth):
    x, y, z = wl.lane_coords(depth)
    bits = ((masks[:, 0][:, None] >> np.arange(8)) & 1).astype(bool).ravel()
    out = dense_yzx.copy()
    out[y[bits], z[bits], x[bits]] = values.reshape(-1)[bits]
    return out


def check_invariants(vx, o, it, roots, depth):
    """refcount == in-degree for every reachable node and nothing else is alive."""
    dl = it.download()
    sig = o.dag_signature(dl["children"], dl["values"], np.array(roots, np.uint64), depth, want_indeg=True)
    live = sig["indeg"] > 0
    assert np.array_equal(dl["refs"][live], sig["indeg"][live])
    assert (dl["refs"][~live] == 0).all()
    st = it.stats()
    assert st["alive_nodes"] - 1 == sig["branches"] + sig["leaves"]
    return sig


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 4, 5])
def test_edit_sequences_match_oracle(gpu_api, oracle_api, depth, dtype):
    vx, o = gpu_api, oracle_api
    rng = np.random.default_rng(100 + depth + 10 * dtype)
    git, oit = vx.VoxInterner.with_memory_budget(64 << 20, dtype), o.VoxInterner(64 << 20, dtype)
    gt, ot = vx.VoxTree(depth, dtype), o.VoxTree(depth, dtype)
    n = 1 << depth
    dense = np.zeros((n, n, n), wl.NP_DTYPE[dtype])
    oracle_ok = True
    for step in range(8):
        masks, values = random_edit(rng, depth, dtype, density=[0.5, 0.05, 0.2, 0.01][step % 4],
                                    alphabet=np.array([1, 2, 3, 7]))
        gb, ob = gt.create_batch(), ot.create_batch()
        set_batch_arrays(gb, masks, values)
        set_batch_arrays(ob, masks, values)
        gchanged = gt.apply_batch(git, gb)
        new_dense = overlay(dense, masks, values, depth)
        assert gchanged == (not np.array_equal(new_dense, dense) or bool(_any_uniform_full(masks, values)))
        dense = new_dense
        assert np.array_equal(gt.to_vec(git), dense)
        sig = check_invariants(vx, o, git, [gt.get_root_id()], depth)
        if oracle_ok:
            try:
                ochanged = ot.apply_batch(oit, ob)
            except o.ReferencePanic:
                oracle_ok = False      # the reference's own latent panic: stop comparing to it
                continue
            assert ochanged == gchanged
            assert np.array_equal(ot.to_vec(oit), dense)
            odl = oit.download()
            osig = o.dag_signature(odl["children"], odl["values"], np.array([ot.get_root_id()], np.uint64), depth,
                                   want_stream=True)
            gdl = git.download()
            gsig = o.dag_signature(gdl["children"], gdl["values"], np.array([gt.get_root_id()], np.uint64), depth,
                                   want_stream=True)
            assert np.array_equal(gsig["stream"], osig["stream"])      # same DAG, same collapse decisions
            assert gsig["per_depth"] == osig["per_depth"]


def _any_uniform_full(masks, values):
    full = masks[:, 0] == 0xFF
    return (full & (values == values[:, :1]).all(1)).any()


def test_alternating_batches_recycle_indices(gpu_api):
    """alternating the two set_sum batches on one tree keeps 177 live nodes; freed
    indices are reused (next_index stops at 355) and carry a bumped generation."""
    vx = gpu_api
    it = vx.VoxInterner.with_memory_budget(24 << 20)
    tree = vx.VoxTree(5)
    b = [tree.create_batch(), tree.create_batch()]
    for k, off in enumerate((1, 100)):
        m, v = wl.batch_from_function(5, wl.p_sum(off), wl.U8, 1)
        set_batch_arrays(b[k], m[0], v[0])
    gens_seen = set()
    for i in range(6):
        assert tree.apply_batch(it, b[i & 1])
        st = it.stats()
        assert st["alive_nodes"] - 1 == 177
        gens_seen.add(vx.api.id_gen(tree.get_root_id()))
        exp = wl.dense_expected(b[i & 1].masks, b[i & 1].values)
        assert np.array_equal(tree.to_vec(it), exp)
    assert it.next_index == 355
    assert it.stats()["recycled_nodes"] == 177
    assert max(gens_seen) >= 1


def test_clear_and_fill_release_everything(gpu_api, oracle_api):
    vx = gpu_api
    it = vx.VoxInterner.with_memory_budget(24 << 20)
    t1, t2 = vx.VoxTree(5), vx.VoxTree(5)
    m, v = wl.batch_from_function(5, wl.p_random(4), wl.U8, 2)
    b1, b2 = t1.create_batch(), t2.create_batch()
    set_batch_arrays(b1, m[0], v[0])
    set_batch_arrays(b2, m[1], v[1])
    assert t1.apply_batch(it, b1) and t2.apply_batch(it, b2)
    alive_both = it.stats()["alive_nodes"]
    t1.clear(it)                                   # voxtree.rs:283-292
    assert t1.is_empty() and t1.is_dirty()
    assert np.array_equal(t2.to_vec(it), wl.dense_expected(m[1], v[1]))     # the other tree is intact
    check_invariants(vx, oracle_api, it, [t2.get_root_id()], 5)
    assert it.stats()["alive_nodes"] < alive_both
    t2.fill(it, 9)                                 # voxtree.rs:264-281: release + Leaf(9)
    assert t2.is_leaf() and it.get_ref(t2.get_root_id()) == 1
    assert it.stats()["alive_nodes"] - 1 == 1
    assert (t2.to_vec(it) == 9).all()
    t2.fill(it, 0)                                 # fill(default) == clear
    assert t2.is_empty() and it.stats()["alive_nodes"] == 1
    # rebuild in the same interner: recycled indices, bumped generations, still correct
    assert t1.apply_batch(it, b1)
    assert np.array_equal(t1.to_vec(it), wl.dense_expected(m[0], v[0]))
    check_invariants(vx, oracle_api, it, [t1.get_root_id()], 5)


def test_apply_batches_mixed_fresh_and_edited(gpu_api, oracle_api):
    vx, o = gpu_api, oracle_api
    rng = np.random.default_rng(3)
    depth = 4
    it = vx.VoxInterner.with_memory_budget(64 << 20)
    trees = [vx.VoxTree(depth) for _ in range(12)]
    dense = [np.zeros((16, 16, 16), np.uint8) for _ in trees]
    for rnd in range(3):
        idx = [i for i in range(12) if (i + rnd) % 3 != 0]     # a different subset every round
        batches = []
        for i in idx:
            masks, values = random_edit(rng, depth, wl.U8, 0.1, np.array([1, 2, 5]))
            b = trees[i].create_batch()
            set_batch_arrays(b, masks, values)
            batches.append(b)
            dense[i] = overlay(dense[i], masks, values, depth)
        vx.apply_batches(it, [trees[i] for i in idx], batches)
        for i, t in enumerate(trees):
            assert np.array_equal(t.to_vec(it), dense[i])
        check_invariants(vx, o, it, [t.get_root_id() for t in trees], depth)


def test_set_root_id_shares_a_tree(gpu_api):
    vx = gpu_api
    it = vx.VoxInterner.with_memory_budget(8 << 20)
    t1, t2 = vx.VoxTree(4), vx.VoxTree(4)
    m, v = wl.batch_from_function(4, wl.p_sum(1), wl.U8, 1)
    b = t1.create_batch()
    set_batch_arrays(b, m[0], v[0])
    assert t1.apply_batch(it, b)
    vx.api._ck(vx.lib().vx_tree_set_root_id(it.h, t2.h, t1.get_root_id()))    # voxtree.rs:135-141
    assert it.get_ref(t1.get_root_id()) == 2
    t1.clear(it)
    assert np.array_equal(t2.to_vec(it), wl.dense_expected(m[0], v[0]))
    t2.clear(it)
    assert it.stats()["alive_nodes"] == 1


# ---------------------------------------------------------------------------------------------- ADVICE (round 1)
@pytest.mark.parametrize("builder", ["fused", "bulk"])
def test_i32_out_of_memory_with_repeated_values_returns(gpu_api, builder, monkeypatch):
    """i32 leaves live in an open-addressing table whose key is claimed before the index is allocated.  When the
    allocation fails ("Out of memory", interner/macros.rs:38) every lane / warp waiting for the same value must be
    told, not left spinning: the call has to come back with VX_E_OOM."""
    vx = gpu_api
    monkeypatch.setenv("VX_BUILDER", builder)
    # ~200 distinct values, each shared by many uniform 2^3 cells (= many lanes ask for the same new leaf at once)
    masks, values = wl.batch_from_function(5, wl.p_random(200, cell=2), wl.I32, 16 if builder == "bulk" else 2)
    g = vx.VoxInterner.with_memory_budget(82 * 48, wl.I32)          # 48 nodes
    with pytest.raises(vx.VoxelisError) as e:
        g.apply_batches_slab(5, masks, values)
    assert e.value.code == -2
    g.reset()
    m2, v2 = wl.batch_from_function(5, wl.p_uniform(-7), wl.I32, 1)
    roots, changed = g.apply_batches_slab(5, m2, v2)
    assert changed[0] == 1 and vx.id_is_leaf(roots[0])


def test_i32_dedup_shard_out_of_memory_returns(gpu_api):
    """The same for the owner side of the global dedup (intern_records_kernel, leaf round)."""
    from voxelis_b200 import dedup
    vx = gpu_api
    masks, values = wl.batch_from_function(4, wl.p_random(200, cell=2), wl.I32, 8)
    locals_, roots = [], []
    for r in range(2):
        it = vx.VoxInterner.with_memory_budget(32 << 20, wl.I32)
        rt, _ = it.apply_batches_slab(4, masks, values)
        locals_.append(it)
        roots.append(rt)
    with pytest.raises(vx.VoxelisError) as e:
        dedup.global_dedup_local(locals_, roots, 82 * 24, wl.I32)   # shards of 24 nodes
    assert e.value.code == -2


def test_i32_leaf_tombstones_are_reclaimed(gpu_api, oracle_api, monkeypatch):
    """Releasing an i32 leaf leaves a tombstone in the leaf table; the periodic table rebuild re-inserts the live
    leaves too, so create / release cycles never eat the table (the reference's map removes the entry,
    interner/mod.rs:276-281).  VX_REHASH_AT forces a rebuild every few releases."""
    vx, o = gpu_api, oracle_api
    monkeypatch.setenv("VX_REHASH_AT", "40")
    g = vx.VoxInterner.with_memory_budget(8 << 20, wl.I32)
    keep = vx.VoxTree(4, wl.I32)
    kb = keep.create_batch()
    km, kv = wl.batch_from_function(4, wl.p_random(9, cell=2), wl.I32, 1)
    kb.assign(km[0], kv[0])
    assert keep.apply_batch(g, kb)
    want_keep = wl.dense_expected(km[0], kv[0])
    for cycle in range(40):
        t = vx.VoxTree(4, wl.I32)
        b = t.create_batch()
        m, v = wl.batch_from_function(4, wl.p_random(60, cell=2), wl.I32, 1, chunk_arg=[1000 + cycle])
        v = (v * 1000 + cycle).astype(np.int32) * (m[:, :, :1] != 0)      # values nobody else uses: fresh leaves every cycle
        b.assign(m[0], v[0])
        assert t.apply_batch(g, b)
        assert np.array_equal(t.to_vec(g), wl.dense_expected(m[0], v[0]))
        t.clear(g)                                                        # releases the cycle's leaves -> tombstones
        assert np.array_equal(keep.to_vec(g), want_keep)                  # the tree that stays is untouched by rebuilds
    st = g.stats()
    c = o.VoxInterner(8 << 20, wl.I32)
    cr, _ = c.apply_batches_fresh(4, km, kv)
    assert st["alive_nodes"] == c.stats()["alive_nodes"]
    # a fresh build after all the churn still finds (not duplicates) the surviving leaves
    roots, _ = g.apply_batches_slab(4, km, kv)
    assert int(roots[0]) == keep.get_root_id() and g.stats()["alive_nodes"] == st["alive_nodes"]


def test_set_root_id_rejects_stale_ids(gpu_api):
    """A released id (free slot / bumped generation), an index never allocated, or a forged generation must be refused
    (reference: assert is_valid_block_id, interner/mod.rs:997-1008) instead of putting a reference on a dead slot."""
    vx = gpu_api
    it = vx.VoxInterner.with_memory_budget(8 << 20)
    t1, t2 = vx.VoxTree(4), vx.VoxTree(4)
    m, v = wl.batch_from_function(4, wl.p_sum(1), wl.U8, 1)
    b = t1.create_batch()
    set_batch_arrays(b, m[0], v[0])
    assert t1.apply_batch(it, b)
    root = t1.get_root_id()
    assert vx.lib().vx_tree_set_root_id(it.h, t2.h, root ^ (1 << 32)) == -1          # wrong generation
    assert vx.lib().vx_tree_set_root_id(it.h, t2.h, (root & ~0xFFFFFFFF) | 60000) == -1   # never allocated
    t1.clear(it)
    assert vx.lib().vx_tree_set_root_id(it.h, t2.h, root) == -1                       # released
    assert t2.is_empty() and it.stats()["alive_nodes"] == 1


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_reset_of_used_state_only_is_a_full_reset(gpu_api, oracle_api, dtype):
    """vx_interner_reset clears only what the nodes in [1, next_index) put into the tables when nothing was ever released
    (reset_used_kernel); every build after such a reset must come out as in a brand-new interner — same DAG, same
    counters, no stale table entry answering a lookup — also after a release (full clear again) and after an overflow."""
    vx, o = gpu_api, oracle_api
    g = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    worlds = [wl.terrain_world((4, 2, 4), 5, "surface_and_below", dtype, materials=3),
              wl.batch_from_function(5, wl.p_random(200, cell=2), dtype, 6),
              wl.named_workload("sum", 3, 5, dtype),
              wl.batch_from_function(5, wl.p_random(4), dtype, 5)]
    for rnd in range(2):
        for m, v in worlds:
            g.reset()
            roots, changed = g.apply_batches_slab(5, m, v)
            c = o.VoxInterner(64 << 20, dtype)
            cr, cc = c.apply_batches_fresh(5, m, v)
            parity.assert_parity(vx, o, 5, g, roots, changed, c, cr, cc)
        # a release in between: the next reset has generations and tombstones to clear
        t = vx.VoxTree(5, dtype)
        b = t.create_batch()
        b.assign(*[a[0] for a in worlds[2]])
        t.apply_batch(g, b)
        t.clear(g)
    # an interner that overflowed is cleared completely as well
    small = vx.VoxInterner.with_memory_budget(82 * 48, dtype)
    with pytest.raises(vx.VoxelisError):
        small.apply_batches_slab(5, *worlds[1])
    small.reset()
    m, v = wl.batch_from_function(5, wl.p_uniform(3), dtype, 1)
    r, ch = small.apply_batches_slab(5, m, v)
    assert ch[0] == 1 and vx.id_is_leaf(r[0]) and small.stats()["alive_nodes"] == 2


@pytest.mark.parametrize("depth,dtype", [(6, wl.U8), (6, wl.I32), (7, wl.U8)], ids=["d6-u8", "d6-i32", "d7-u8"])
def test_single_batch_apply_on_big_trees(gpu_api, depth, dtype):
    """vx_tree_apply_batch on one handle (apply_one_in_place, vx_capi.cu): a D = 6 u8 batch crosses the bus as one copy,
    a D = 6 i32 batch and a D = 7 batch are read in place from the pinned arena (D = 7: by the bulk builder's plan
    kernel, 32-byte vector loads over the bus).  Fresh build, an edit on top of it, a fill, every voxel checked."""
    vx = gpu_api
    rng = np.random.default_rng(depth * 10 + dtype)
    it = vx.VoxInterner.with_memory_budget(1 << 30, dtype)
    t = vx.VoxTree(depth, dtype)
    n = 1 << depth
    dense = np.zeros((n, n, n), wl.NP_DTYPE[dtype])
    for step, density in enumerate((0.3, 0.02)):
        masks, values = random_edit(rng, depth, dtype, density, np.array([1, 2, 3, 7]), uniform_blocks=0.3)
        b = t.create_batch()
        if step == 0:
            b.assign(masks, values)            # API-written batch (journal / occupancy kept)
        else:
            set_batch_arrays(b, masks, values)  # raw arrays
        assert t.apply_batch(it, b)
        dense = overlay(dense, masks, values, depth)
        assert np.array_equal(t.to_vec(it), dense)
    b = t.create_batch()
    b.fill(it, 9)
    assert t.apply_batch(it, b) and t.is_leaf()
    assert (t.to_vec(it) == 9).all()

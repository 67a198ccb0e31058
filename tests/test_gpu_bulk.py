"""The two builders for fresh trees — the fused apply_kernel and the level-synchronous bulk pipeline
(voxelis_b200/csrc/vx_bulk.cuh) — against the oracle and against each other.  VX_BUILDER pins one."""
import numpy as np
import pytest

import parity
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _mixed(depth, dtype, n_small):
    parts = [wl.batch_from_function(depth, wl.p_random(4), dtype, n_small),
             wl.batch_from_function(depth, wl.p_random(255, cell=2), dtype, 3),
             wl.batch_from_function(depth, wl.p_sparse(), dtype, 4),
             wl.named_workload("checkerboard", 5, depth, dtype),
             wl.named_workload("hollow", 2, depth, dtype),
             wl.named_workload("uniform", 3, depth, dtype)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    masks[1] = 0                       # nothing set: the tree stays EMPTY, changed == false
    masks[2, ::3, 0] = 0               # ragged: whole blocks without set bits
    masks[4, 100:, 0] = 0              # only the first 100 blocks carry patches
    return masks, values


@pytest.mark.parametrize("builder", ["bulk", "fused"])
@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [4, 5, 6])
def test_builders_match_oracle(gpu_api, oracle_api, depth, dtype, builder, monkeypatch):
    monkeypatch.setenv("VX_BUILDER", builder)
    masks, values = _mixed(depth, dtype, 24 if depth < 6 else 3)
    r = parity.build_both(gpu_api, oracle_api, depth, masks, values, dtype, budget=512 << 20)
    assert r[2][1] == 0 and r[1][1] == 0
    parity.assert_parity(gpu_api, oracle_api, depth, *r)


def test_bulk_depth7(gpu_api, oracle_api, monkeypatch):
    monkeypatch.setenv("VX_BUILDER", "bulk")
    parts = [wl.batch_from_function(7, wl.p_random(255, cell=4), wl.U8, 1),
             wl.batch_from_function(7, wl.p_sparse(3, 8), wl.U8, 2)]
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    r = parity.build_both(gpu_api, oracle_api, 7, masks, values, wl.U8, budget=512 << 20)
    parity.assert_parity(gpu_api, oracle_api, 7, *r, dense_check=False)


def test_bulk_terrain_and_repeat_calls(gpu_api, oracle_api, monkeypatch):
    """Two calls into one interner (the second finds the first one's nodes), slices forced by a tiny scratch
    limit, and the terrain world: same DAG as the oracle's serial application."""
    vx, o = gpu_api, oracle_api
    monkeypatch.setenv("VX_BUILDER", "bulk")
    monkeypatch.setenv("VX_BULK_MAX_BYTES", str(3 << 20))   # ~50 chunks of 32^3 per slice
    m1, v1 = wl.terrain_world((6, 3, 6), 5, "surface_and_below", wl.U8, materials=3)
    m2, v2 = wl.terrain_world((5, 3, 5), 5, "surface_only", wl.U8, x_chunk_offset=3)
    g = vx.VoxInterner.with_memory_budget(256 << 20)
    r1, c1 = g.apply_batches_slab(5, m1, v1)
    r2, c2 = g.apply_batches_slab(5, m2, v2)
    c = o.VoxInterner(256 << 20)
    cr, cc = c.apply_batches_fresh(5, np.concatenate([m1, m2]), np.concatenate([v1, v2]))
    parity.assert_parity(vx, o, 5, g, np.concatenate([r1, r2]), np.concatenate([c1, c2]), c, cr, cc)


def test_bulk_after_release_uses_free_list(gpu_api, oracle_api, monkeypatch):
    """Nodes released by VoxTree::clear go to the free list (interner/macros.rs:1-41); a bulk build after
    that recycles them with bumped generations, exactly like the fused kernel does."""
    vx, o = gpu_api, oracle_api
    m, v = wl.batch_from_function(5, wl.p_random(255, cell=2), wl.U8, 6)
    results = {}
    for builder in ("fused", "bulk"):
        monkeypatch.setenv("VX_BUILDER", builder)
        g = vx.VoxInterner.with_memory_budget(256 << 20)
        t = vx.VoxTree(5)
        b = t.create_batch()
        b.masks[:] = m[0]
        b.values[:] = v[0]
        b.mark_patched()
        t.apply_batch(g, b)
        before = g.stats()["alive_nodes"]
        t.clear(g)                                   # everything but the empty branch is released
        assert g.stats()["alive_nodes"] == 1 and before > 500
        roots, changed = g.apply_batches_slab(5, m, v)
        st = g.stats()
        dense = g.roots_to_vec(roots, 5)
        for i in range(len(roots)):
            assert np.array_equal(dense[i], wl.dense_expected(m[i], v[i]))
        gens = np.array([(int(r) >> 32) & 0x7FFF for r in roots])
        results[builder] = (st["alive_nodes"], st["recycled_nodes"], int(g.next_index), sorted(gens.tolist()))
        d = g.download()
        sig = o.dag_signature(d["children"], d["values"], roots, 5, want_stream=True, want_indeg=True)
        live = sig["numbers"] != 0
        assert np.array_equal(d["refs"][live], sig["indeg"][live])
        results[builder + "_sig"] = sig["sig"]
    assert results["fused"][:3] == results["bulk"][:3]
    assert results["fused_sig"] == results["bulk_sig"]


def test_bulk_host_paths_and_device_path_agree(gpu_api, monkeypatch):
    import torch
    vx = gpu_api
    monkeypatch.setenv("VX_BUILDER", "bulk")
    masks, values = wl.terrain_world((8, 2, 8), 5, "surface_only", wl.U8)
    ref = None
    for mode in ("pageable", "pinned", "pinned_staged", "device"):
        g = vx.VoxInterner.with_memory_budget(128 << 20)
        if mode == "pageable":
            roots, changed = g.apply_batches_slab(5, masks, values)
        elif mode.startswith("pinned"):
            if mode == "pinned_staged":
                monkeypatch.setenv("VX_HOST_MODE", "staged")
            hm, hv = torch.from_numpy(masks).pin_memory(), torch.from_numpy(values).pin_memory()
            roots, changed = g.apply_batches_slab(5, hm.numpy(), hv.numpy())
            monkeypatch.delenv("VX_HOST_MODE", raising=False)
        else:
            dm, dv = torch.from_numpy(masks).cuda(), torch.from_numpy(values).cuda()
            dr = torch.zeros(len(masks), dtype=torch.int64, device="cuda")
            dc = torch.zeros(len(masks), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            g.apply_batches_device(5, len(masks), dm.data_ptr(), dv.data_ptr(), dr.data_ptr(), dc.data_ptr())
            g.sync()
            roots, changed = dr.cpu().numpy().view(np.uint64), dc.cpu().numpy()
        st = g.stats()
        dense = g.roots_to_vec(roots[:16], 5)
        got = (st["alive_nodes"], st["leaf_nodes"], st["collapsed_branches"], changed.tobytes(), dense.tobytes())
        ref = ref or got
        assert got == ref, mode


def test_bulk_out_of_memory_is_a_status(gpu_api, monkeypatch):
    vx = gpu_api
    monkeypatch.setenv("VX_BUILDER", "bulk")
    masks, values = wl.batch_from_function(5, wl.p_random(255), wl.U8, 16)
    g = vx.VoxInterner.with_memory_budget(1 << 20)            # ~13k nodes; the batch needs ~75k
    with pytest.raises(vx.VoxelisError) as e:
        g.apply_batches_slab(5, masks, values)
    assert "Out of memory" in str(e.value)
    with pytest.raises(vx.VoxelisError):
        g.apply_batches_slab(5, masks[:1], values[:1])        # poisoned until reset
    g.reset()
    roots, changed = g.apply_batches_slab(5, *wl.named_workload("checkerboard", 4, 5, wl.U8))
    assert changed.all()


def test_bulk_repeated_calls_varied_sizes(gpu_api, monkeypatch):
    """The bulk scratch (level lists, epoch-tagged flags) is reused across calls of any size and depth:
    every call must come out as if it were the only one."""
    vx = gpu_api
    monkeypatch.setenv("VX_BUILDER", "bulk")
    g = vx.VoxInterner.with_memory_budget(512 << 20)
    rng = np.random.default_rng(5)
    world = wl.terrain_world((8, 3, 8), 5, "surface_and_below", wl.U8, materials=2)
    for step, (depth, n) in enumerate([(5, 64), (5, 192), (6, 3), (5, 7), (4, 40), (5, 192), (7, 1), (5, 64)]):
        if depth == 5:
            pick = rng.choice(world[0].shape[0], n, replace=False)
            masks, values = world[0][pick].copy(), world[1][pick].copy()
            masks[rng.integers(0, n)] = 0
        else:
            masks, values = wl.batch_from_function(depth, wl.p_random(255, cell=4 if depth > 5 else 2), wl.U8, n)
            masks[n // 2, ::2, 0] = 0
        roots, changed = g.apply_batches_slab(depth, masks, values)
        assert np.array_equal(changed.astype(bool), masks[:, :, 0].any(1)), step
        assert ((roots == 0) == (changed == 0)).all(), step
        check = rng.choice(n, min(n, 6), replace=False)
        dense = g.roots_to_vec(roots[check], depth)
        for j, i in enumerate(check):
            assert np.array_equal(dense[j], wl.dense_expected(masks[i], values[i])), (step, i)


def test_bulk_lod_values_are_stable(gpu_api, oracle_api, monkeypatch):
    """Regression: the second level of bulk_upper_kernel read the values of children created in the SAME
    launch through L1 and could hit a line cached before another SM wrote the value (rare wrong LOD value,
    about one run in three at D = 6).  Same input many times: the branch LOD values must equal the oracle's
    every time."""
    vx, o = gpu_api, oracle_api
    monkeypatch.setenv("VX_BUILDER", "bulk")
    masks, values = _mixed(6, wl.U8, 3)
    c = o.VoxInterner(512 << 20, wl.U8)
    flags, fills = parity.flags_from(masks.shape[0])
    croots, cchanged = c.apply_batches_fresh(6, masks, values, flags & 1, fills, (flags >> 1) & 1)
    cd = c.download()
    cs = o.dag_signature(cd["children"], cd["values"], croots, 6, want_stream=True)
    corder = np.argsort(cs["numbers"], kind="stable")[np.count_nonzero(cs["numbers"] == 0):]
    g = vx.VoxInterner.with_memory_budget(512 << 20, wl.U8)
    for rep in range(12):
        g.reset()
        groots, _ = g.apply_batches_slab(6, masks, values)
        gd = g.download()
        gs = o.dag_signature(gd["children"], gd["values"], groots, 6, want_stream=True)
        gorder = np.argsort(gs["numbers"], kind="stable")[np.count_nonzero(gs["numbers"] == 0):]
        assert np.array_equal(gs["stream"], cs["stream"]), rep
        assert np.array_equal(gd["values"][gorder], cd["values"][corder]), rep


# ---------------------------------------------------------------------------------------------- unit memo
def _repeating_world(dtype):
    """Busy units whose content repeats: identical chunks (checkerboard, sum), 255-periodic chunks, solid rock next
    to them, and unique high-entropy chunks in between (every unit of those is its own representative)."""
    parts = [wl.named_workload("checkerboard", 40, 5, dtype),
             wl.named_workload("sum", 24, 5, dtype),
             wl.batch_from_function(5, wl.p_random(255), dtype, 3),
             wl.named_workload("sum_per_chunk", 300, 5, dtype),      # chunk c and c + 255 are identical
             wl.named_workload("uniform", 6, 5, dtype),
             wl.named_workload("checkerboard", 9, 5, dtype),
             wl.terrain_world((3, 2, 3), 5, "surface_and_below", dtype, materials=3)]
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("memo", ["1", "0"])
def test_unit_memo_matches_oracle(gpu_api, oracle_api, dtype, memo, monkeypatch):
    """Aliased units take the node AND the counters of the first unit with their content: DAG, refcounts and every
    InternerStats counter equal the oracle's serial application, with the memo on and off."""
    monkeypatch.setenv("VX_BUILDER", "bulk")
    monkeypatch.setenv("VX_UNIT_MEMO", memo)
    masks, values = _repeating_world(dtype)
    r = parity.build_both(gpu_api, oracle_api, 5, masks, values, dtype, budget=256 << 20)
    parity.assert_parity(gpu_api, oracle_api, 5, *r)


def test_unit_memo_is_per_call(gpu_api, oracle_api, monkeypatch):
    """Entries of one call never serve another (they name units of that call): three calls into one interner with
    overlapping content, a reset in between, and a call large enough to trigger the table wipe."""
    vx, o = gpu_api, oracle_api
    monkeypatch.setenv("VX_BUILDER", "bulk")
    g = vx.VoxInterner.with_memory_budget(2 << 30)
    c = o.VoxInterner(2 << 30)
    m1, v1 = wl.named_workload("checkerboard", 20, 5)
    m2, v2 = wl.named_workload("sum", 20, 5)
    m3, v3 = wl.batch_from_function(5, wl.p_random(255), wl.U8, 4200)     # 33 600 distinct busy units: beyond MEMO_CLEAR_AT
    groots, croots, gch, cch = [], [], [], []
    for m, v in ((m1, v1), (m2, v2), (np.concatenate([m2, m1]), np.concatenate([v2, v1])), (m3[:16], v3[:16]), (m1, v1)):
        r, ch = g.apply_batches_slab(5, m, v)
        groots.append(r)
        gch.append(ch)
        r, ch = c.apply_batches_fresh(5, m, v)
        croots.append(r)
        cch.append(ch)
    parity.assert_parity(vx, o, 5, g, np.concatenate(groots), np.concatenate(gch), c, np.concatenate(croots), np.concatenate(cch))
    # the big call: wipes the table at its end; the calls after it still alias correctly
    g.reset()
    r3, _ = g.apply_batches_slab(5, m3, v3)
    ra, _ = g.apply_batches_slab(5, np.concatenate([m1, m2]), np.concatenate([v1, v2]))
    rb, _ = g.apply_batches_slab(5, m2, v2)
    assert len(set(ra[:20].tolist())) == 1 and len(set(ra[20:].tolist())) == 1 and ra[20] == rb[0]
    dense = g.roots_to_vec(np.concatenate([r3[:4], ra[:1], rb[:1]]), 5)
    for i, (m, v) in enumerate([(m3[0], v3[0]), (m3[1], v3[1]), (m3[2], v3[2]), (m3[3], v3[3]), (m1[0], v1[0]), (m2[0], v2[0])]):
        assert np.array_equal(dense[i], wl.dense_expected(m, v))
    st = g.stats()
    assert st["branch_nodes"] - 1 == 4200 * 4681 + 5 + 83       # all-miss chunks + checkerboard + sum


@pytest.mark.parametrize("name", ["checkerboard", "sum"])
def test_config2_full_size_counters_and_dag(gpu_api, oracle_api, name):
    """BASELINE config 2 at its full size (4 096 chunks of 32^3 into one interner, voxtree_bench.rs:699-710,777-786):
    with 32 767 of the 32 768 units aliased through the unit memo, every InternerStats counter, the DAG and the
    refcounts still equal the oracle's serial application, and every chunk gets the same root."""
    vx, o = gpu_api, oracle_api
    masks, values = wl.named_workload(name, 4096)
    r = parity.build_both(vx, o, 5, masks, values, wl.U8, budget=256 << 20)
    g, groots = r[0], r[1]
    assert len(set(groots.tolist())) == 1
    parity.assert_parity(vx, o, 5, *r, dense_check=False)
    st = g.stats()
    want = {"checkerboard": (5, 1), "sum": (83, 94)}[name]
    assert st["branch_nodes"] - 1 == want[0] and st["leaf_nodes"] == want[1]          # SURVEY §8c known answers
    assert np.array_equal(g.roots_to_vec(groots[-1:], 5)[0], wl.dense_expected(masks[-1], values[-1]))

"""Shared parity checks: CUDA path (through the C ABI) vs the CPU oracle on identical batches."""
import numpy as np

from voxelis_b200 import workloads as wl


def flags_from(n, fill=None, has_patches=True):
    flags = np.full(n, (1 if fill is not None else 0) | (2 if has_patches else 0), np.uint8)
    fills = np.full(n, 0 if fill is None else fill, np.int64)
    return flags, fills


def build_both(vx, o, depth, masks, values, dtype=0, fill=None, has_patches=True, budget=64 << 20):
    n = masks.shape[0]
    flags, fills = flags_from(n, fill, has_patches)
    g = vx.VoxInterner.with_memory_budget(budget, dtype)
    groots, gchanged = g.apply_batches_slab(depth, masks, values, flags, fills)
    c = o.VoxInterner(budget, dtype)
    croots, cchanged = c.apply_batches_fresh(depth, masks, values, flags & 1, fills, (flags >> 1) & 1)
    return g, groots, gchanged, c, croots, cchanged


def assert_parity(vx, o, depth, g, groots, gchanged, c, croots, cchanged, dense_check=True, stats=True,
                  refs=True, indeg=True):
    """Bit-exact voxels, isomorphic DAG (identical canonical record stream), identical per-depth
    unique counts, collapse decisions, hit/miss counters, refcounts and LOD values."""
    assert np.array_equal(gchanged, cchanged)
    gd, cd = g.download(), c.download()
    gs = o.dag_signature(gd["children"], gd["values"], groots, depth, want_stream=True, want_indeg=True)
    cs = o.dag_signature(cd["children"], cd["values"], croots, depth, want_stream=True, want_indeg=True)
    assert gs["per_depth"] == cs["per_depth"]
    assert (gs["branches"], gs["leaves"]) == (cs["branches"], cs["leaves"])
    assert np.array_equal(gs["stream"], cs["stream"])          # isomorphic up to node-id permutation
    assert gs["sig"] == cs["sig"]
    # same number of live nodes; no holes in the index space
    assert gd["n"] == cd["n"]
    # per-node payload compared in canonical order
    gn, cn = gs["numbers"], cs["numbers"]
    gorder = np.argsort(gn, kind="stable")[np.count_nonzero(gn == 0):]
    corder = np.argsort(cn, kind="stable")[np.count_nonzero(cn == 0):]
    assert np.array_equal(gd["values"][gorder], cd["values"][corder])      # leaf values + branch LOD values
    if refs:
        assert np.array_equal(gd["refs"][gorder], cd["refs"][corder])
        if indeg:   # in-degree invariant (not with a fill: phase 0 leaks one reference)
            assert np.array_equal(gd["refs"][gorder], gs["indeg"][gorder])
    # types/mask/leaf bits of every stored child id agree with what the child is
    ch = gd["children"][gorder]
    for node_children in ch[:2048]:
        for cid in node_children:
            cid = int(cid)
            if cid:
                assert (cid & 0xFFFFFFFF) < gd["n"]
    if stats:
        gst, cst = g.stats(), c.stats()
        for k in ("collapsed_branches", "leaf_nodes", "branch_nodes", "total_cache_hits", "total_cache_misses",
                  "leaf_cache_hits", "leaf_cache_misses", "branch_cache_hits", "branch_cache_misses",
                  "alive_nodes", "allocated_nodes", "nodes_capacity", "node_size"):
            assert gst[k] == cst[k], (k, gst[k], cst[k])
    if dense_check:
        gdense = g.roots_to_vec(groots, depth)
        for i in range(len(groots)):
            assert np.array_equal(gdense[i], c.root_to_vec(croots[i], depth))
    return gs

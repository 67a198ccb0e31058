"""Cross-checks the oracle restatement against the independent canonical builder
(tests/canonical.py) and re-derives the known answers of SURVEY.md §8(c)."""
import numpy as np
import pytest

from canonical import CanonicalDag
from voxelis_b200 import workloads as wl

PATTERNS = {
    "uniform": wl.p_uniform(1),
    "uniform_half": wl.p_uniform_half(1),
    "checkerboard_bench": wl.p_checkerboard_bench(1),
    "checkerboard_test": wl.p_checkerboard_test(),
    "sum": wl.p_sum(1),
    "sparse": wl.p_sparse(),
    "hollow": wl.p_hollow_cube(),
    "diagonal": wl.p_diagonal(),
    "gradient": wl.p_gradient(),
    "random255": wl.p_random(255),
    "random4": wl.p_random(4),
    "cell4": wl.p_random(255, cell=4),
}


def oracle_build(o, depth, masks, values, dtype=0, fill=None, budget=64 << 20):
    it = o.VoxInterner(budget, dtype)
    n = masks.shape[0]
    hf = None if fill is None else np.ones(n, np.uint8)
    fv = None if fill is None else np.full(n, fill, np.int64)
    roots, changed = it.apply_batches_fresh(depth, masks, values, hf, fv)
    return it, roots, changed


def check_against_canonical(o, it, roots, depth, masks, values, fill=None, check_refs=True):
    dl = it.download()
    sig = o.dag_signature(dl["children"], dl["values"], roots, depth, want_stream=True, want_indeg=True)
    dag = CanonicalDag()
    croots = []
    for c in range(masks.shape[0]):
        dense = wl.dense_expected(masks[c], values[c], fill)
        assert np.array_equal(it.root_to_vec(roots[c], depth), dense)
        croots.append(dag.build(dense))
    assert np.array_equal(sig["stream"], dag.stream(croots))
    assert sig["per_depth"] == dag.per_depth(croots, depth)
    assert (sig["branches"], sig["leaves"]) == dag.reachable(croots)
    if check_refs:
        # refcount(n) = in-edges from unique live branches + root handles
        live = sig["indeg"] > 0
        assert np.array_equal(dl["refs"][live], sig["indeg"][live])
        assert (dl["refs"][~live] == 0).all()
        # every allocated node is reachable: no garbage after a fresh build
        assert live[1:].all()
    return sig


@pytest.mark.parametrize("depth", [2, 3, 4, 5])
@pytest.mark.parametrize("name", sorted(PATTERNS))
def test_patterns_vs_canonical(oracle_api, name, depth):
    masks, values = wl.batch_from_function(depth, PATTERNS[name], wl.U8, 1)
    it, roots, changed = oracle_build(oracle_api, depth, masks, values)
    assert changed[0] == (masks[0, :, 0].any())
    check_against_canonical(oracle_api, it, roots, depth, masks, values)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32])
def test_multi_chunk_shared_interner(oracle_api, dtype):
    masks, values = wl.batch_from_function(4, wl.p_random(4), dtype, 6)
    m2, v2 = wl.batch_from_function(4, wl.p_sum_per_chunk(), dtype, 3)
    masks, values = np.concatenate([masks, m2, masks[:2]]), np.concatenate([values, v2, values[:2]])
    it, roots, _ = oracle_build(oracle_api, 4, masks, values, dtype)
    assert roots[0] == roots[9] and roots[1] == roots[10]      # dedup across trees
    check_against_canonical(oracle_api, it, roots, 4, masks, values)


def test_fill_with_patches_vs_canonical(oracle_api):
    rng = np.random.default_rng(5)
    for depth in (3, 4):
        B = wl.blocks_per_chunk(depth)
        masks = np.zeros((4, B, 2), np.uint8)
        values = np.zeros((4, B, 8), np.uint8)
        sel = rng.random((4, B, 8)) < 0.05
        vals = rng.integers(1, 4, (4, B, 8)).astype(np.uint8)
        vals[vals == 3] = 7          # never equal to the fill value 3 (SURVEY §0 refcount quirk)
        values[sel] = vals[sel]
        masks[:, :, 0] = (sel * (1 << np.arange(8))).sum(-1).astype(np.uint8)
        it, roots, changed = oracle_build(oracle_api, depth, masks, values, fill=3)
        assert changed.all()
        # phase-0 leaves one extra reference on the fill leaf per apply: skip refs
        check_against_canonical(oracle_api, it, roots, depth, masks, values, fill=3, check_refs=False)


KNOWN = {  # fresh 32^3 chunk -> branches, leaves, collapsed, (hits, misses), per-depth
    "uniform": (0, 1, 4681, (4095, 1), None),
    "uniform_half": (1, 1, 2340, (2047, 2), [(1, 0), (0, 1)]),
    "checkerboard_bench": (5, 1, 0, (21059, 6), [(1, 0)] * 5 + [(0, 1)]),
    "sum": (83, 94, 0, (37272, 177), [(1, 0), (4, 0), (10, 0), (22, 0), (46, 0), (0, 94)]),
    "sparse": (5, 1, 0, (1603, 6), [(1, 0)] * 5 + [(0, 1)]),
    "hollow": (87, 1, 0, (7393, 88), [(1, 0), (8, 0), (26, 0), (26, 0), (26, 0), (0, 1)]),
    "diagonal": (5, 1, 0, (57, 6), [(1, 0)] * 5 + [(0, 1)]),
}


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_known_answers_d5(oracle_api, name):
    branches, leaves, collapsed, (hits, misses), per_depth = KNOWN[name]
    masks, values = wl.batch_from_function(5, PATTERNS[name], wl.U8, 1)
    it, roots, _ = oracle_build(oracle_api, 5, masks, values)
    st = it.stats()
    assert st["branch_nodes"] - 1 == branches        # slot 0 (empty branch) is counted, mod.rs:131
    assert st["leaf_nodes"] == leaves
    assert st["collapsed_branches"] == collapsed
    assert (st["total_cache_hits"], st["total_cache_misses"]) == (hits, misses)
    dl = it.download()
    sig = oracle_api.dag_signature(dl["children"], dl["values"], roots, 5)
    assert (sig["branches"], sig["leaves"]) == (branches, leaves)
    if per_depth is not None:
        got = [p for p in sig["per_depth"] if p != (0, 0)]
        assert got == per_depth
    else:
        assert oracle_api.id_is_leaf(int(roots[0]))


def test_alternating_batches_recycle(oracle_api):
    """alternating the two set_sum batches on one tree keeps 177 live nodes and
    stops next_index at 355 (free-list LIFO, interner/macros.rs:1-41)."""
    o = oracle_api
    it = o.VoxInterner(24 << 20)
    tree = o.VoxTree(5)
    b = [tree.create_batch(), tree.create_batch()]
    for k, off in enumerate((1, 100)):
        m, v = wl.batch_from_function(5, wl.p_sum(off), wl.U8, 1)
        b[k].masks[:], b[k].values[:], b[k].has_patches = m[0], v[0], True
    for i in range(6):
        assert tree.apply_batch(it, b[i & 1])
        st = it.stats()
        assert st["alive_nodes"] - 1 == 177
    assert it.next_index == 355
    assert it.free_count == 177


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_to_vec_at_every_lod_matches_numpy_pyramid(oracle_api, dtype):
    """to_vec(interner, root, max_depth.for_lod(lod)) (world/voxchunk.rs:267): branches at the cut-off depth
    contribute their LOD value = calc_average of their children (core/voxel.rs:96-141)."""
    from canonical import lod_pyramid
    o = oracle_api
    for name, pat in (("random4", wl.p_random(4)), ("cell2", wl.p_random(255, cell=2)), ("sparse", wl.p_sparse()),
                      ("hollow", wl.p_hollow_cube()), ("uniform", wl.p_uniform(3))):
        depth = 4
        masks, values = wl.batch_from_function(depth, pat, dtype, 1)
        c = o.VoxInterner(64 << 20, dtype)
        roots, _ = c.apply_batches_fresh(depth, masks, values)
        pyr = lod_pyramid(wl.dense_expected(masks[0], values[0]))
        for lod in range(depth + 2):
            got = c.root_to_vec(int(roots[0]), depth, lod)
            want = pyr[min(lod, depth)]
            assert np.array_equal(got, want), (name, lod)

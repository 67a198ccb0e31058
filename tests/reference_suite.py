"""The reference's own batch-path tests, restated once and run against BOTH the CPU oracle
(tests/test_oracle_reference_tests.py) and the CUDA path through the C ABI
(tests/test_gpu_reference_tests.py).

Each function cites the Rust test it follows in /root/reference/voxelis/src/spatial/voxtree.rs.
``api`` is a module exposing VoxInterner / VoxTree / Batch with the reference's method names
(oracle.oracle or voxelis_b200).
"""
import itertools
import random

import numpy as np

D5 = 5
BUDGET = 1024 * 1024


def _all_positions(n):
    return itertools.product(range(n), range(n), range(n))


def _check_every_voxel(api, interner, tree, expect_fn):
    """assert tree.get(pos) == expect_fn(x,y,z) for every voxel, through get() itself and
    through to_vec (utils/common.rs:158-246)."""
    n = tree.voxels_per_axis()
    exp = np.zeros((n, n, n), np.int64)  # [y][z][x]
    for y, z, x in _all_positions(n):
        v = expect_fn(x, y, z)
        exp[y, z, x] = 0 if v is None else v
    if hasattr(tree, "get_many"):
        xs, ys, zs = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        pos = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], 1).astype(np.int32)
        found, vals = tree.get_many(interner, pos)
        got = np.where(found.astype(bool), vals, 0)
        want = exp[pos[:, 1], pos[:, 2], pos[:, 0]]
        assert np.array_equal(found.astype(bool), want != 0)
        assert np.array_equal(got, want)
        # and the scalar entry point on a sample
        rng = random.Random(1234)
        for _ in range(64):
            x, y, z = (rng.randrange(n) for _ in range(3))
            assert tree.get(interner, (x, y, z)) == expect_fn(x, y, z)
    else:
        for y, z, x in _all_positions(n):
            assert tree.get(interner, (x, y, z)) == expect_fn(x, y, z), (x, y, z)
    dense = tree.to_vec(interner)
    assert np.array_equal(dense.astype(np.int64), exp)


def batch_double_apply(api):
    """voxtree.rs:1345-1362 test_batch_double_apply"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    batch = tree.create_batch()
    batch.set(interner, (0, 0, 0), 3)
    assert tree.apply_batch(interner, batch) is True
    assert tree.apply_batch(interner, batch) is False


def batch_expand_shared_leaf(api):
    """voxtree.rs:1399-1434 test_patterns_batch_expand_shared_leaf"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    batch = tree.create_batch()
    batch.fill(interner, 1)
    batch.set(interner, (0, 0, 0), 2)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: 2 if (x, y, z) == (0, 0, 0) else 1)
    assert not tree.is_empty()
    assert not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_checkerboard(api):
    """voxtree.rs:1474-1511 test_patterns_batch_checkerboard"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    batch = tree.create_batch()
    for y, z, x in _all_positions(n):
        assert batch.set(interner, (x, y, z), 2 if (x + y + z) % 2 == 0 else 1)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: 2 if (x + y + z) % 2 == 0 else 1)
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_solid_fill_one_by_one(api):
    """voxtree.rs:1546-1580 test_patterns_batch_solid_fill_one_by_one"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    batch = tree.create_batch()
    for y, z, x in _all_positions(n):
        assert batch.set(interner, (x, y, z), 3)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: 3)
    assert not tree.is_empty()
    assert tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_solid_fill_half_one_by_one(api):
    """voxtree.rs:1627-1673 test_patterns_batch_solid_fill_half_one_by_one — five successive
    batches on the SAME tree (exercises apply on a non-empty tree + release of the old one)."""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    half = n // 2
    for value in range(1, 6):
        batch = tree.create_batch()
        for y in range(half):
            for z in range(n):
                for x in range(n):
                    assert batch.set(interner, (x, y, z), value)
        assert tree.apply_batch(interner, batch)
        _check_every_voxel(api, interner, tree, lambda x, y, z: value if y < half else None)
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_solid_fill_fill_op(api):
    """voxtree.rs:1701-1729 test_patterns_batch_solid_fill_fill_op"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    batch = tree.create_batch()
    batch.fill(interner, 3)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: 3)
    assert not tree.is_empty()
    assert tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_sparse_fill(api):
    """voxtree.rs:1770-1810 test_patterns_batch_sparse_fill (every 4th voxel)"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    batch = tree.create_batch()
    for y in range(0, n, 4):
        for z in range(0, n, 4):
            for x in range(0, n, 4):
                assert batch.set(interner, (x, y, z), 3)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree,
                       lambda x, y, z: 3 if (x % 4 == 0 and y % 4 == 0 and z % 4 == 0) else None)
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_gradient_fill(api):
    """voxtree.rs:1853-1892 test_patterns_batch_gradient_fill (value = x % 256; x = 0 is a clear)"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    batch = tree.create_batch()
    for x in range(n):
        for y in range(n):
            for z in range(n):
                assert batch.set(interner, (x, y, z), x % 256)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: (x % 256) or None)
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_hollow_cube(api):
    """voxtree.rs:1950-2005 test_patterns_batch_hollow_cube"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()

    def edge(x, y, z):
        return x in (0, n - 1) or y in (0, n - 1) or z in (0, n - 1)

    batch = tree.create_batch()
    for y, z, x in _all_positions(n):
        if edge(x, y, z):
            assert batch.set(interner, (x, y, z), 3)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: 3 if edge(x, y, z) else None)
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_diagonal(api):
    """voxtree.rs:2044-2082 test_patterns_batch_diagonal"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    batch = tree.create_batch()
    for i in range(n):
        assert batch.set(interner, (i, i, i), 3)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: 3 if x == y == z else None)
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def batch_random_noise(api, seed=7):
    """voxtree.rs:2134-2185 test_patterns_batch_random_noise (1000 random (pos, value))"""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    tree = api.VoxTree(D5)
    n = tree.voxels_per_axis()
    rng = random.Random(seed)
    data = {}
    batch = tree.create_batch()
    for _ in range(1000):
        p = (rng.randrange(n), rng.randrange(n), rng.randrange(n))
        v = rng.randint(1, 255)
        data[p] = v
        assert batch.set(interner, p, v)
    assert tree.apply_batch(interner, batch)
    _check_every_voxel(api, interner, tree, lambda x, y, z: data.get((x, y, z)))
    assert not tree.is_empty() and not tree.is_leaf()
    assert interner.get_ref(tree.get_root_id()) == 1


def shared_interner_deduplication(api):
    """voxtree.rs:1284-1321 (uniqueness + deduplication), restated on the batch path: two
    trees in one interner — same content => same root id; different content => different."""
    interner = api.VoxInterner.with_memory_budget(BUDGET)
    t1, t2, t3 = api.VoxTree(3), api.VoxTree(3), api.VoxTree(3)
    assert t1.is_empty() and t2.is_empty()
    b = t1.create_batch()
    b.set(interner, (0, 0, 0), 42)
    assert t1.apply_batch(interner, b)
    assert t1.get(interner, (0, 0, 0)) == 42
    assert t2.get(interner, (0, 0, 0)) is None
    b2 = t2.create_batch()
    b2.set(interner, (0, 0, 0), 42)
    assert t2.apply_batch(interner, b2)
    assert t1.get_root_id() == t2.get_root_id()
    assert interner.get_ref(t1.get_root_id()) == 2
    b3 = t3.create_batch()
    b3.set(interner, (0, 0, 0), 24)
    assert t3.apply_batch(interner, b3)
    assert t3.get(interner, (0, 0, 0)) == 24
    assert t3.get_root_id() != t1.get_root_id()


ALL = [batch_double_apply, batch_expand_shared_leaf, batch_checkerboard, batch_solid_fill_one_by_one,
       batch_solid_fill_half_one_by_one, batch_solid_fill_fill_op, batch_sparse_fill,
       batch_gradient_fill, batch_hollow_cube, batch_diagonal, batch_random_noise,
       shared_interner_deduplication]

"""Pins the CPU oracle (oracle/voxelis_oracle.hpp) against every known-answer the
reference's own test-suite holds for the batch path."""
import numpy as np
import pytest

import reference_suite as rs


@pytest.mark.parametrize("case", rs.ALL, ids=lambda f: f.__name__)
def test_reference_batch_tests(oracle_api, case):
    case(oracle_api)


def test_block_id_kats(oracle_api):
    """core/block_id.rs:411 doctest + :474-541 unit tests."""
    o = oracle_api
    assert o.id_pack(123, 456, 0, 0, True) == 0x800001C80000007B           # doctest :411
    leaf = o.id_pack(123, 456, 0, 0, True)
    assert o.id_index(leaf) == 123 and o.id_gen(leaf) == 456 and o.id_is_leaf(leaf)
    br = o.id_pack(123, 456, 0xAB, 0xCD, False)
    assert (o.id_index(br), o.id_gen(br), o.id_types(br), o.id_mask(br)) == (123, 456, 0xAB, 0xCD)
    assert o.id_is_branch(br)
    assert o.id_pack(0, 0, 0, 0, False) == 0                                # EMPTY, :112 / test_empty
    mx = o.id_pack(0xFFFFFFFF, 0x7FFE, 0xFF, 0xFF, True)                   # test_max_values
    assert mx != 0xFFFFFFFFFFFFFFFF


def test_max_depth_limits(oracle_api):
    """core/max_depth.rs:167-172: MaxDepth::new(7) panics in the reference.  Depth 7 is this
    build's documented extension; 8 must be rejected."""
    with pytest.raises(oracle_api.ReferencePanic):
        oracle_api.VoxTree(8)
    assert oracle_api.VoxTree(6).voxels_per_axis() == 64


def test_path_masks_match_reference_table(oracle_api):
    """spatial/voxtree.rs:41-105 rows, retyped here as the known answer for the formula."""
    L = oracle_api.lib()
    table = {
        1: [0b111],
        2: [0b111_000, 0b111_111],
        3: [0b111_000_000, 0b111_111_000, 0b111_111_111],
        4: [0b111 << 9, 0b111_111 << 6, 0b111_111_111 << 3, 0b111_111_111_111],
        5: [0b111 << 12, 0b111_111 << 9, 0b111_111_111 << 6, 0b111_111_111_111 << 3, (1 << 15) - 1],
        6: [0b111 << 15, 0b111_111 << 12, 0b111_111_111 << 9, 0b111_111_111_111 << 6,
            ((1 << 15) - 1) << 3, (1 << 18) - 1],
    }
    for d, row in table.items():
        for lvl, want in enumerate(row):
            assert L.orc_path_mask(d, lvl) == want
        for lvl in range(len(row), 6):
            assert L.orc_path_mask(d, lvl) == 0
    for lvl in range(6):
        assert L.orc_path_mask(0, lvl) == 0


def test_morton_path(oracle_api):
    """utils/common.rs:24-55: x -> bit 3n, y -> 3n+1, z -> 3n+2."""
    L = oracle_api.lib()
    assert L.orc_encode_child_index_path(1, 0, 0) == 1
    assert L.orc_encode_child_index_path(0, 1, 0) == 2
    assert L.orc_encode_child_index_path(0, 0, 1) == 4
    assert L.orc_encode_child_index_path(2, 0, 0) == 8
    assert L.orc_encode_child_index_path(31, 31, 31) == (1 << 15) - 1
    assert L.orc_encode_child_index_path(1023, 1023, 1023) == (1 << 30) - 1


def test_budget_and_capacity(oracle_api):
    """interner/mod.rs:49-68,158-164: capacity = budget / (78 + sizeof T)."""
    o = oracle_api
    assert o.VoxInterner(256 << 20, o.U8).capacity == (256 << 20) // 79 == 3397917
    assert o.VoxInterner(1 << 20, o.I32).capacity == (1 << 20) // 82
    with pytest.raises(o.ReferencePanic):
        o.VoxInterner(10, o.U8)                     # "Requested budget is too small"
    small = o.VoxInterner(79 * 4, o.U8)             # 4 nodes: index 0 reserved -> 3 usable
    t = o.VoxTree(3)
    b = t.create_batch()
    for i, v in enumerate((1, 2, 3, 4)):
        b.set(small, (i, 0, 0), v)
    with pytest.raises(o.ReferencePanic, match="Out of memory"):  # interner/macros.rs:38
        t.apply_batch(small, b)


def test_fill_plus_equal_patches_quirk(oracle_api):
    """fill(F) + patches all equal to F drops the whole batch (returns false)."""
    o = oracle_api
    it = o.VoxInterner(BUDGET := 1 << 20)
    t = o.VoxTree(5)
    b = t.create_batch()
    b.fill(it, 3)
    b.set(it, (1, 2, 3), 3)
    assert t.apply_batch(it, b) is False
    assert t.is_empty()


def test_clear_inside_batch_is_ignored(oracle_api):
    """SURVEY §0 / voxtree.rs:778-781: clear_mask is never read."""
    o = oracle_api
    it = o.VoxInterner(1 << 20)
    t = o.VoxTree(4)
    b = t.create_batch()
    b.set(it, (1, 1, 1), 5)
    b.set(it, (1, 1, 1), 0)      # cancels the set, records a clear
    b.set(it, (2, 2, 2), 7)
    assert t.apply_batch(it, b)
    assert t.get(it, (1, 1, 1)) is None
    assert t.get(it, (2, 2, 2)) == 7

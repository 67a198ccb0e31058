"""Reference pin of the occupancy planes: tests/golden/occupancy_known_answer.py writes out, by hand from
utils/mesh.rs:263-285,418-596, the plane words of three small worlds.  The oracle's restatement and the CUDA path
(vx_occupancy_masks) must both produce exactly those words."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import occupancy_known_answer as ka  # noqa: E402

from voxelis_b200 import workloads as wl  # noqa: E402


def batches(vols, dtype):
    ms, vs = zip(*[wl.batch_from_dense(v, v != 0, dtype) for v, _ in vols])
    return np.stack(ms), np.stack(vs), [o for _, o in vols]


def check(got, want, nm):
    assert list(got["material_ids"][:nm]) == want["material_ids"]
    assert list(got["material_counts"][:nm]) == want["material_counts"]
    assert np.array_equal(got["global"], want["global"])
    assert np.array_equal(got["active"], want["active"])
    for k in range(nm):
        assert np.array_equal(got["per_material"][k], want["per_material"][k])


def test_literals_agree_with_the_written_out_loops():
    a = ka.case_a()
    for idx, word in ka.CASE_A_LITERALS.items():
        assert int(a["global"][idx]) == word
    assert int(np.count_nonzero(a["global"])) == ka.CASE_A_NONZERO_WORDS
    b = ka.case_b()
    for idx, word in ka.CASE_B_LITERALS_GLOBAL.items():
        assert int(b["global"][idx]) == word
    assert [int(x) for x in b["active"]] == ka.CASE_B_ACTIVE


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_oracle_matches_hand_derivation(oracle_api, dtype):
    o = oracle_api
    # case A
    m, v, offs = batches([ka.case_a_volume()], dtype)
    c = o.VoxInterner(16 << 20, dtype)
    roots, _ = c.apply_batches_fresh(5, m, v)
    assert c.stats()["leaf_nodes"] == 1                       # the cube really is ONE leaf (of side 4)
    check(c.occupancy_masks(roots, 5, offs), ka.case_a(), 1)
    # case B
    m, v, offs = batches(ka.case_b_volumes(), dtype)
    c = o.VoxInterner(16 << 20, dtype)
    roots, _ = c.apply_batches_fresh(5, m, v)
    check(c.occupancy_masks(roots, 5, offs), ka.case_b(), 2)
    # case C
    m, v = wl.batch_from_function(6, wl.p_uniform(5), dtype, 1)
    c = o.VoxInterner(16 << 20, dtype)
    roots, _ = c.apply_batches_fresh(6, m, v)
    got = c.occupancy_masks(roots, 6, [(0, 0, 0)])
    assert list(got["material_ids"]) == ka.CASE_C["material_ids"] and list(got["material_counts"]) == ka.CASE_C["material_counts"]
    assert (got["global"] == np.uint64(ka.CASE_C["word"])).all() and (got["active"] == np.uint64(ka.CASE_C["word"])).all()
    assert (got["per_material"][0] == np.uint64(ka.CASE_C["word"])).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_cuda_matches_hand_derivation(gpu_api, dtype):
    vx = gpu_api
    for vols, want, nm, depth in (([ka.case_a_volume()], ka.case_a(), 1, 5), (ka.case_b_volumes(), ka.case_b(), 2, 5)):
        m, v, offs = batches(vols, dtype)
        g = vx.VoxInterner.with_memory_budget(16 << 20, dtype)
        roots, _ = g.apply_batches_slab(depth, m, v)
        for M in (nm, 8):                                     # shared-memory kernel and the word-owner kernels' sizing
            got = g.occupancy_masks(roots, depth, offs, max_materials=M)
            assert got["n_materials"][0] == nm
            one = {"material_ids": got["material_ids"][0], "material_counts": got["material_counts"][0], "global": got["global"][0],
                   "active": got["active"][0], "per_material": got["per_material"][0]}
            check(one, want, nm)
    m, v = wl.batch_from_function(6, wl.p_uniform(5), dtype, 1)
    g = vx.VoxInterner.with_memory_budget(16 << 20, dtype)
    roots, _ = g.apply_batches_slab(6, m, v)
    got = g.occupancy_masks(roots, 6, [(0, 0, 0)], max_materials=2)
    assert got["n_materials"][0] == 1 and got["material_ids"][0][0] == 5 and got["material_counts"][0][0] == 262144
    assert (got["global"][0] == np.uint64(ka.CASE_C["word"])).all() and (got["active"][0] == np.uint64(ka.CASE_C["word"])).all()
    assert (got["per_material"][0][0] == np.uint64(ka.CASE_C["word"])).all()

"""Reference pin of the terrain batch rules: tests/golden/terrain_known_answer.py derives the Batch
arrays of one 8^3 chunk by hand from utils/shapes.rs:273-357 + core/batch.rs:145-175.  The numpy generator that feeds
every other test and the bench (workloads.terrain_world) and the CUDA generator (vx_terrain_batches_device) must both
produce exactly those arrays."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import terrain_known_answer as ka  # noqa: E402

from voxelis_b200 import workloads as wl  # noqa: E402

CASES = [(True, 1), (False, 1), (False, 3), (True, 3)]


def test_literals_agree_with_the_written_out_rules():
    m, v = ka.expected(True, 1)
    for p, i, val in ka.LITERAL_SURFACE:
        assert m[p, 0] >> i & 1 and v[p, i] == val
    for p, i in ka.LITERAL_SURFACE_ABSENT:
        assert not (m[p, 0] >> i & 1)
    assert int(np.unpackbits(m[:, 0]).sum()) == 64            # one voxel per column
    m, v = ka.expected(False, 3)
    for p, i, val in ka.LITERAL_3MAT:
        assert m[p, 0] >> i & 1 and v[p, i] == val, (p, i, val, v[p])
    assert int(np.unpackbits(m[:, 0]).sum()) == 62 + 8 + 5    # 62 columns of height 0, one of 7, one of 4


@pytest.mark.parametrize("surface_only,materials", CASES)
@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_numpy_generator_matches_hand_derivation(monkeypatch, surface_only, materials, dtype):
    monkeypatch.setattr(wl, "height_field", lambda nx, nz, seed=0, height=8, x0=0, z0=0: ka.HEIGHTS.copy())
    m, v = wl.terrain_world((1, 1, 1), 3, "surface_only" if surface_only else "surface_and_below", dtype, materials=materials)
    em, ev = ka.expected(surface_only, materials, wl.NP_DTYPE[dtype])
    assert np.array_equal(m[0][:, 0], em[:, 0]) and np.array_equal(v[0], ev)


@pytest.mark.gpu
@pytest.mark.parametrize("surface_only,materials", CASES)
@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_cuda_generator_matches_hand_derivation(gpu_api, surface_only, materials, dtype):
    import torch
    g = gpu_api.VoxInterner.with_memory_budget(16 << 20, dtype)
    h = torch.from_numpy(ka.HEIGHTS.copy()).cuda()
    dm = torch.zeros((1, 64, 2), dtype=torch.uint8, device="cuda")
    dv = torch.zeros((1, 64, 8), dtype=torch.uint8 if dtype == wl.U8 else torch.int32, device="cuda")
    torch.cuda.synchronize()
    g.terrain_batches_device(3, (1, 1, 1), h.data_ptr(), dm.data_ptr(), dv.data_ptr(), surface_only, materials)
    g.sync()
    em, ev = ka.expected(surface_only, materials, wl.NP_DTYPE[dtype])
    assert np.array_equal(dm.cpu().numpy()[0][:, 0], em[:, 0])
    assert np.array_equal(dv.cpu().numpy()[0], ev)

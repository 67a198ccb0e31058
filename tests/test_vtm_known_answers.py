"""Reference pin of the VTM writer / reader: the bytes in tests/golden/vtm_known_answer.py are derived
by hand from world/voxmodel.rs:177-283, world/voxchunk.rs:382-405, io/varint.rs:5-17 and io/export.rs:90-151.  The
oracle must write exactly them (its pool indices follow the reference's allocation order); the CUDA path must import
them, re-export them byte for byte, and write them itself up to the order in which its lanes created the two leaves."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import vtm_known_answer as ka  # noqa: E402

from voxelis_b200 import workloads as wl  # noqa: E402

POS = np.array([[1, -2, 3]], np.int32)


def world(dtype, half=False):
    n = 4
    x, y, z = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    vol = np.where(z < 2, 1, 0 if half else 2)
    return wl.batch_from_dense(vol, None, dtype)


def expected_dense(half):
    out = np.zeros((4, 4, 4), np.int64)            # [y][z][x]
    out[:, :2, :] = 1
    if not half:
        out[:, 2:, :] = 2
    return out


@pytest.mark.parametrize("dtype,vb", [(wl.U8, 1), (wl.I32, 4)], ids=["u8", "i32"])
@pytest.mark.parametrize("half", [False, True])
def test_oracle_writes_the_hand_derived_bytes(oracle_api, dtype, vb, half):
    m, v = world(dtype, half)
    c = oracle_api.VoxInterner(1 << 20, dtype)
    roots, _ = c.apply_batches_fresh(2, m[None], v[None])
    want = ka.payload_half(vb) if half else ka.payload(vb)
    assert c.model_serialize(POS, roots) == want
    # and reads them: node k of the file becomes pool index k (interner/mod.rs:933,948)
    c2 = oracle_api.VoxInterner(1 << 20, dtype)
    pos, r2 = c2.model_deserialize(want)
    assert np.array_equal(pos, POS)
    assert np.array_equal(c2.root_to_vec(int(r2[0]), 2), expected_dense(half))
    assert (int(r2[0]) & 0xFFFFFFFF) == (2 if half else 3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,vb", [(wl.U8, 1), (wl.I32, 4)], ids=["u8", "i32"])
@pytest.mark.parametrize("half", [False, True])
def test_gpu_reads_and_writes_the_hand_derived_bytes(gpu_api, dtype, vb, half, tmp_path):
    vx = gpu_api
    want = ka.payload_half(vb) if half else ka.payload(vb)
    # import: ids of the file are pool indices -> the re-export is the same bytes
    g = vx.VoxInterner.with_memory_budget(1 << 20, dtype)
    pos, roots = g.model_deserialize(want)
    assert np.array_equal(pos, POS) and (int(roots[0]) & 0xFFFFFFFF) == (2 if half else 3)
    assert np.array_equal(g.roots_to_vec(roots, 2)[0], expected_dense(half))
    assert g.model_serialize(pos, roots) == want
    # build + export: the two leaves are created by different lanes of one warp step, so their pool order is the
    # hardware's; either order is the reference's file up to that swap
    m, v = world(dtype, half)
    g2 = vx.VoxInterner.with_memory_budget(1 << 20, dtype)
    r2, _ = g2.apply_batches_slab(2, m[None], v[None])
    got = g2.model_serialize(POS, r2)
    val = lambda x: int(x).to_bytes(vb, "big")
    swapped = want if half else (ka.be32(2) + b"\x01" + val(2) + b"\x02" + val(1) + ka.be32(1) + b"\x03\xff" + b"\x02" * 4 + b"\x01" * 4 +
                                 val(1) + want[-29:])
    assert got in (want, swapped)
    # the whole file, uncompressed (io/export.rs:90-151)
    path = str(tmp_path / "pin.vtm")
    g.export_vtm(path, "pin", 2, 1.25, (4, 5, 6), pos, roots, compress=False)
    assert open(path, "rb").read() == ka.file_bytes(want, hashlib.md5(want).digest())


def test_file_header_layout():
    """The header arithmetic of golden/vtm_known_answer.file_bytes against the reader used by the other VTM tests."""
    import struct
    import vtm_ref
    p = ka.PAYLOAD_U8
    raw = ka.file_bytes(p, hashlib.md5(p).digest())
    assert raw[:12] == b"VoxTreeModel" and struct.unpack(">f", raw[17:21])[0] == 1.25
    assert len(raw) == 12 + 2 + 2 + 1 + 4 + 8 + 12 + 1 + 3 + 16 + 4 + len(p)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".vtm", delete=False) as f:
        f.write(raw)
    info = vtm_ref.read_vtm(f.name)
    os.unlink(f.name)
    assert info["payload"] == p and info["name"] == "pin" and info["world_bounds"] == (4, 5, 6) and info["max_depth"] == 2

"""VTM export from the device pools: VoxModel::serialize (world/voxmodel.rs:177-294) and
export_model_to_vtm (io/export.rs:90-151).

* byte-exact against a plain-Python restatement of the writer applied to the pools downloaded from the GPU;
* the payload read back by the oracle's restatement of the reference IMPORTER (voxmodel.rs:296-408) gives the
  same voxels in every chunk, the same node counts, and refcount == in-degree (+1 per root);
* same size as the oracle's own serialisation of the oracle-built world (ids are a permutation, so equal
  record sizes are only guaranteed while every id fits one varint byte class — checked on a small world);
* file: header fields, MD5 of the payload, zstd level-7 stream when libzstd is present."""
import os

import numpy as np
import pytest

import vtm_ref
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def build(vx, o, depth, dtype):
    parts = [wl.batch_from_function(depth, wl.p_random(4), dtype, 3),
             wl.named_workload("sum", 2, depth, dtype),
             wl.named_workload("uniform", 1, depth, dtype),
             wl.named_workload("hollow", 1, depth, dtype)]
    if depth == 5:
        parts.append(wl.terrain_world((3, 2, 3), 5, "surface_and_below", dtype, materials=3))
    masks = np.concatenate([p[0] for p in parts])
    values = np.concatenate([p[1] for p in parts])
    masks[1] = 0                                   # an EMPTY chunk: its root maps to id 0
    n = masks.shape[0]
    g = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    groots, _ = g.apply_batches_slab(depth, masks, values)
    c = o.VoxInterner(64 << 20, dtype)
    croots, _ = c.apply_batches_fresh(depth, masks, values)
    positions = np.stack([np.arange(n) % 5 - 2, np.arange(n) // 5, -np.arange(n)], 1).astype(np.int32)
    return g, groots, c, croots, positions


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5])
def test_payload_is_exact_and_imports(gpu_api, oracle_api, depth, dtype):
    vx, o = gpu_api, oracle_api
    g, groots, c, croots, positions = build(vx, o, depth, dtype)
    payload = g.model_serialize(positions, groots)
    d = g.download()
    expect = vtm_ref.payload_from_pools(d["children"], d["values"], d["refs"], 1 if dtype == wl.U8 else 4, positions, groots)
    assert payload == expect
    # through the restated reference importer into a fresh CPU interner
    c2 = o.VoxInterner(64 << 20, dtype)
    pos2, roots2 = c2.model_deserialize(payload)
    assert np.array_equal(pos2, positions)
    dense = g.roots_to_vec(groots, depth)
    for i in range(len(groots)):
        assert np.array_equal(dense[i], c2.root_to_vec(int(roots2[i]), depth)), i
    assert c2.stats()["alive_nodes"] == g.stats()["alive_nodes"]
    d2 = c2.download()
    sig = o.dag_signature(d2["children"], d2["values"], roots2, depth, want_indeg=True)
    assert np.array_equal(d2["refs"][1:], sig["indeg"][1:])
    # and the oracle's own file of the oracle-built world holds the same number of records
    ref_payload = c.model_serialize(positions, croots)
    assert ref_payload[:4] == payload[:4]                       # leaf count
    if g.stats()["alive_nodes"] < 128:                          # every id is a one-byte varint: sizes must agree
        assert len(ref_payload) == len(payload)


def test_after_edits_and_release(gpu_api, oracle_api):
    """Recycled slots (refcount 0, stale rows cleared) must not be written."""
    vx, o = gpu_api, oracle_api
    depth = 4
    g = vx.VoxInterner.with_memory_budget(32 << 20)
    trees = [vx.VoxTree(depth) for _ in range(4)]
    rng = np.random.default_rng(1)
    for rnd in range(3):
        batches = [t.create_batch() for t in trees]
        for b in batches:
            k = 200
            b.set_many(rng.integers(0, 16, (k, 3)), rng.integers(1, 5, k))
        vx.apply_batches(g, trees, batches)
    trees[3].clear(g)
    roots = np.array([t.get_root_id() for t in trees], np.uint64)
    positions = np.arange(12, dtype=np.int32).reshape(4, 3)
    payload = g.model_serialize(positions, roots)
    d = g.download()
    assert payload == vtm_ref.payload_from_pools(d["children"], d["values"], d["refs"], 1, positions, roots)
    c2 = o.VoxInterner(32 << 20)
    _, roots2 = c2.model_deserialize(payload)
    for t, r in zip(trees, roots2):
        assert np.array_equal(t.to_vec(g), c2.root_to_vec(int(r), depth))


@pytest.mark.parametrize("compress", [False, True], ids=["raw", "zstd"])
def test_vtm_file(gpu_api, oracle_api, tmp_path, compress):
    vx, o = gpu_api, oracle_api
    g, groots, c, croots, positions = build(vx, o, 5, wl.U8)
    path = os.path.join(tmp_path, "world.vtm")
    decompress = None
    if compress:
        pa = pytest.importorskip("pyarrow")
        if not pa.Codec.is_available("zstd"):
            pytest.skip("no zstd decoder to check the stream with")
        payload_len = len(g.model_serialize(positions, groots))
        decompress = lambda b: pa.Codec("zstd").decompress(b, decompressed_size=payload_len).to_pybytes()
    try:
        g.export_vtm(path, "dunes", 5, 1.28, (3, 2, 3), positions, groots, compress=compress)
    except vx.VoxelisError as e:
        if compress and e.code == -4:
            pytest.skip("libzstd not present on this box")
        raise
    f = vtm_ref.read_vtm(path, decompress)
    assert f["flags"] == (1 if compress else 0) and f["max_depth"] == 5 and f["name"] == "dunes"
    assert abs(f["chunk_world_size"] - 1.28) < 1e-6 and f["world_bounds"] == (3, 2, 3) and f["reserved"] == (0, 0)
    assert f["payload"] == g.model_serialize(positions, groots)
    c2 = o.VoxInterner(64 << 20)
    _, roots2 = c2.model_deserialize(f["payload"])
    assert np.array_equal(g.roots_to_vec(groots[:3], 5)[2], c2.root_to_vec(int(roots2[2]), 5))


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_import_installs_a_working_interner(gpu_api, oracle_api, dtype):
    """vx_model_deserialize: same voxels, refcount == in-degree, and the hash tables really hold the imported
    nodes — further edits on the imported trees find them (same live-node count as the oracle doing the same
    on its own import of the same payload)."""
    vx, o = gpu_api, oracle_api
    depth = 4
    g, groots, c, croots, positions = build(vx, o, depth, dtype)
    payload = g.model_serialize(positions, groots)
    g2, c2 = vx.VoxInterner.with_memory_budget(64 << 20, dtype), o.VoxInterner(64 << 20, dtype)
    pos_g, roots_g = g2.model_deserialize(payload)
    pos_c, roots_c = c2.model_deserialize(payload)
    assert np.array_equal(pos_g, positions) and np.array_equal(pos_c, positions)
    assert np.array_equal(roots_g, roots_c)                     # file id == pool index on both sides
    assert np.array_equal(g2.roots_to_vec(roots_g, depth), g.roots_to_vec(groots, depth))
    d2, dc = g2.download(), c2.download()
    assert d2["n"] == dc["n"]
    assert np.array_equal(d2["children"], dc["children"]) and np.array_equal(d2["values"], dc["values"])
    assert np.array_equal(d2["refs"][1:], dc["refs"][1:])
    # a second interner state: export of the import is the same file
    assert g2.model_serialize(pos_g, roots_g) == payload
    # edits on top: the batch of chunk 0 applied to every imported tree
    masks, values = wl.batch_from_function(depth, wl.p_random(4), dtype, 1)
    # (each imported root already holds its tree's reference, voxtree.rs:135-141: the handles adopt it)
    b = vx.Batch(depth, dtype)
    b.assign(masks[0], values[0])
    gtrees = [vx.VoxTree(depth, dtype) for _ in roots_g]
    for t, rg in zip(gtrees, roots_g):
        t.adopt_root(int(rg))
    vx.apply_batches(g2, gtrees, [b] * len(gtrees))
    ob = o.Batch(depth, dtype)
    ob.masks[:], ob.values[:], ob.has_patches = masks[0], values[0], True
    for i, rc in enumerate(roots_c):
        t = o.VoxTree(depth, dtype)
        t.adopt_root(int(rc))
        t.apply_batch(c2, ob)
        assert np.array_equal(gtrees[i].to_vec(g2), t.to_vec(c2)), i
    assert g2.stats()["alive_nodes"] == c2.stats()["alive_nodes"]


def test_import_vtm_file_and_errors(gpu_api, oracle_api, tmp_path):
    vx, o = gpu_api, oracle_api
    g, groots, c, croots, positions = build(vx, o, 5, wl.U8)
    for compress in (False, True):
        path = os.path.join(tmp_path, f"w{int(compress)}.vtm")
        try:
            g.export_vtm(path, "dunes", 5, 2.5, (3, 2, 3), positions, groots, compress=compress)
        except vx.VoxelisError as e:
            if compress and e.code == -4:
                continue
            raise
        g2 = vx.VoxInterner.with_memory_budget(64 << 20)
        meta, pos, roots = g2.import_vtm(path)
        assert meta == {"flags": int(compress), "max_depth": 5, "chunk_world_size": 2.5, "world_bounds": (3, 2, 3),
                        "name": "dunes"}
        assert np.array_equal(pos, positions)
        assert np.array_equal(g2.roots_to_vec(roots, 5), g.roots_to_vec(groots, 5))
        # not fresh any more: a second import is refused (the reference would trip its index assertion)
        with pytest.raises(vx.VoxelisError):
            g2.import_vtm(path)
    # a corrupted payload fails the MD5 check before anything is installed
    raw = bytearray(open(os.path.join(tmp_path, "w0.vtm"), "rb").read())
    raw[-5] ^= 0x40
    bad = os.path.join(tmp_path, "bad.vtm")
    open(bad, "wb").write(raw)
    g3 = vx.VoxInterner.with_memory_budget(64 << 20)
    with pytest.raises(vx.VoxelisError):
        g3.import_vtm(bad)
    assert g3.stats()["alive_nodes"] == 1
    with pytest.raises(vx.VoxelisError):
        g3.model_deserialize(g.model_serialize(positions, groots)[:-3])      # truncated
    assert g3.stats()["alive_nodes"] == 1


def test_import_rejects_cycles_and_duplicate_nodes(gpu_api):
    """A payload whose graph is not a DAG of distinct nodes is refused before the interner is touched (the reference
    trusts the file and would loop or panic; a GPU walk over a cycle never ends)."""
    import struct
    vx = gpu_api
    be = lambda v: struct.pack(">I", v)
    chunk = b"VoxTreeChunk" + be(0) + be(0) + be(0)
    # ids: 1 = Leaf(1); 2, 3 = branches.  (a) branch 2 -> child 3, branch 3 -> child 2: a cycle
    cyc = be(1) + b"\x01\x01" + be(2) + b"\x02\x01\x03\x01" + b"\x03\x01\x02\x01" + be(1) + chunk + b"\x02"
    # (b) branch 2 lists itself
    selfref = be(1) + b"\x01\x01" + be(1) + b"\x02\x03\x01\x02\x01" + be(1) + chunk + b"\x02"
    # (c) two leaves with one value
    dup_leaf = be(2) + b"\x01\x05" + b"\x02\x05" + be(0) + be(1) + chunk + b"\x01"
    # (d) two branches with the same children
    dup_branch = be(1) + b"\x01\x01" + be(2) + b"\x02\x01\x01\x01" + b"\x03\x01\x01\x01" + be(1) + chunk + b"\x03"
    for name, data in (("cycle", cyc), ("self", selfref), ("dup leaf", dup_leaf), ("dup branch", dup_branch)):
        g = vx.VoxInterner.with_memory_budget(1 << 20)
        with pytest.raises(vx.VoxelisError) as e:
            g.model_deserialize(data)
        assert e.value.code == -1, name
        assert g.next_index == 1, name                       # nothing was installed
    # the well-formed sibling of (d) goes through
    ok = be(1) + b"\x01\x01" + be(2) + b"\x02\x01\x01\x01" + b"\x03\x03\x01\x02\x01" + be(1) + chunk + b"\x03"
    g = vx.VoxInterner.with_memory_budget(1 << 20)
    pos, roots = g.model_deserialize(ok)
    assert len(roots) == 1 and g.next_index == 4

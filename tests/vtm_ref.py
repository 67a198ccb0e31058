"""Plain-Python restatement of the VTM writer/reader for tests (reference: world/voxmodel.rs:177-294,
world/voxchunk.rs:382-405, io/export.rs:90-151, io/import.rs:14-98, io/varint.rs)."""
import hashlib
import struct

import numpy as np


def varint(v: int) -> bytes:                       # io/varint.rs:5-32
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def payload_from_pools(children, values, refs, value_bytes, positions, roots) -> bytes:
    """VoxModel::serialize applied to downloaded pools: alive nodes (refcount > 0) renumbered leaves first,
    then branches, each in index order (:199-215); records :231-268; chunk table in the order given."""
    n = len(refs)
    alive = refs > 0
    alive[0] = False
    is_branch = alive & (children != 0).any(axis=1)
    is_leaf = alive & ~is_branch
    newid = np.zeros(n, np.int64)
    leaves = np.nonzero(is_leaf)[0]
    branches = np.nonzero(is_branch)[0]
    newid[leaves] = 1 + np.arange(len(leaves))
    newid[branches] = 1 + len(leaves) + np.arange(len(branches))
    be = lambda v: int(v).to_bytes(value_bytes, "big", signed=value_bytes > 1)
    out = bytearray(struct.pack(">I", len(leaves)))
    for i in leaves:
        out += varint(int(newid[i])) + be(values[i])
    out += struct.pack(">I", len(branches))
    for i in branches:
        row = children[i]
        mask = sum(1 << k for k in range(8) if row[k] != 0)
        out += varint(int(newid[i])) + bytes([mask])
        for k in range(8):
            if row[k] != 0:
                out += varint(int(newid[int(row[k]) & 0xFFFFFFFF]))
        out += be(values[i])
    out += struct.pack(">I", len(roots))
    for p, r in zip(positions, roots):
        out += b"VoxTreeChunk" + struct.pack(">iii", *[int(x) for x in p])
        out += varint(int(newid[int(r) & 0xFFFFFFFF]) if int(r) else 0)
    return bytes(out)


def read_vtm(path, decompress=None):
    """import_model_from_vtm's header walk (io/import.rs:25-88) -> dict; checks magic, version and the MD5."""
    raw = open(path, "rb").read()
    assert raw[:12] == b"VoxTreeModel"
    version, flags = struct.unpack(">HH", raw[12:16])
    assert version == 0x0100
    depth = raw[16]
    (chunk_world_size,) = struct.unpack(">f", raw[17:21])
    r1, r2 = struct.unpack(">II", raw[21:29])
    bounds = struct.unpack(">iii", raw[29:41])
    name_len = raw[41]
    name = raw[42:42 + name_len].decode()
    at = 42 + name_len
    md5 = raw[at:at + 16]
    (size,) = struct.unpack(">I", raw[at + 16:at + 20])
    data = raw[at + 20:at + 20 + size]
    assert len(data) == size and at + 20 + size == len(raw)
    if flags & 1:
        assert decompress is not None, "compressed VTM needs a zstd decompressor"
        data = decompress(data)
    assert hashlib.md5(data).digest() == md5
    return {"flags": flags, "max_depth": depth, "chunk_world_size": chunk_world_size, "reserved": (r1, r2),
            "world_bounds": bounds, "name": name, "payload": data}

"""Hand-derived VTM bytes (reference pin for SURVEY §8f-2).  Nothing here was produced by this repository's code:
every byte is written out from the reference's writer, line by line.  Paths under /root/reference/voxelis/src.

The model: ONE chunk at position (1, -2, 3), MaxDepth 2 (4^3 voxels = 8 blocks of 2^3), built on a fresh interner by
a batch that sets every voxel with z < 2 to 1 and every voxel with z >= 2 to 2.

What the reference builds (spatial/voxtree.rs:724-1118, interner/mod.rs):
  * slot 0 is the permanent empty branch (interner/mod.rs:101-131) and sits in the BRANCH pattern map.
  * Phase 1 walks batch.masks() in index order (voxtree.rs:778).  Block p = Morton(x>>1, y>>1, z>>1) = bx | by<<1 | bz<<2
    (utils/common.rs:24-55), so blocks 0..3 have z < 2 and blocks 4..7 have z >= 2.  Every block has set_mask 0xFF and
    eight equal values -> uniform collapse to a leaf (:826): block 0 creates Leaf(1) = pool index 1 (get_next_index,
    interner/macros.rs:1-41: next_index starts at 1), blocks 1..3 hit it, block 4 creates Leaf(2) = pool index 2.
  * Phase 2 joins the eight blocks into the root: children = [L1, L1, L1, L1, L2, L2, L2, L2]; not eight identical
    ids -> no collapse (:1050) -> get_or_create_branch = pool index 3, types = mask = 0xFF,
    LOD value = calc_average([1,1,1,1,2,2,2,2]) (core/voxel.rs:96-141): counts 4 / 4, the first maximum stays unless it
    is the default value -> 1.

VoxModel::serialize (world/voxmodel.rs:177-283):
  :199-215  id_map: leaves sorted by pool index get 1, 2; branches sorted by pool index: index 0 is skipped, index 3 -> 3
"""

def be32(v):
    return int(v & 0xFFFFFFFF).to_bytes(4, "big")


def payload(value_bytes: int) -> bytes:
    val = lambda v: int(v).to_bytes(value_bytes, "big")        # ByteConversion::write_as_be, core/voxel.rs:26-28
    out = b""
    out += be32(2)                      # :226       writer.write_u32::<BigEndian>(leaf_size)          2 leaves
    out += b"\x01" + val(1)             # :227-235   encode_varint_u32(1) = 01 (io/varint.rs:5-17), value 1 big-endian
    out += b"\x02" + val(2)             #            encode_varint_u32(2) = 02, value 2
    out += be32(1)                      # :237       write_u32(branch_size - 1): the map holds slot 0 and the root
    out += b"\x03"                      # :243-247   new id of the root, varint
    out += b"\xff"                      # :248       id.mask(): all eight children present
    out += b"\x01\x01\x01\x01"          # :250-261   children 0..3 -> new id 1 (Leaf(1))
    out += b"\x02\x02\x02\x02"          #            children 4..7 -> new id 2 (Leaf(2))
    out += val(1)                       # :262-263   branch LOD value, big-endian
    out += be32(1)                      # :276-279   number of chunks
    out += b"VoxTreeChunk"              # world/voxchunk.rs:392   VTC_MAGIC (io/consts.rs:3)
    out += be32(1) + be32(-2) + be32(3)  # :396-398  position x, y, z as big-endian i32
    out += b"\x03"                      # :400-404   encode_varint(new id of the root)
    return out


PAYLOAD_U8 = payload(1)
PAYLOAD_I32 = payload(4)
assert PAYLOAD_U8.hex() == ("00000002" "0101" "0202" "00000001" "03" "ff" "0101010102020202" "01" "00000001"
                            "566f7854726565436875 6e6b".replace(" ", "") + "00000001" "fffffffe" "00000003" "03")
assert len(PAYLOAD_U8) == 52 and len(PAYLOAD_I32) == 61

# The same world with only the z < 2 half set: children 4..7 are EMPTY -> mask 0x0F, four child ids, and
# calc_average([1,1,1,1,0,0,0,0]) = 1 (a tie never moves from a non-default value to the default one).
def payload_half(value_bytes: int) -> bytes:
    val = lambda v: int(v).to_bytes(value_bytes, "big")
    return (be32(1) + b"\x01" + val(1) + be32(1) + b"\x02" + b"\x0f" + b"\x01\x01\x01\x01" + val(1) +
            be32(1) + b"VoxTreeChunk" + be32(1) + be32(-2) + be32(3) + b"\x02")


def file_bytes(payload_: bytes, md5_digest: bytes, name=b"pin", depth=2, chunk_world_size_be=b"\x3f\xa0\x00\x00",
               bounds=(4, 5, 6)) -> bytes:
    """export_model_to_vtm (io/export.rs:90-151) with Flags::NONE (the zstd stream of Flags::DEFAULT is not something a
    person derives by hand; the flag only switches the payload's encoding, :133-139)."""
    out = b"VoxTreeModel"               # :108  VTM_MAGIC (io/consts.rs:2)
    out += b"\x01\x00"                  # :109  VTM_VERSION 0x0100, big-endian u16
    out += b"\x00\x00"                  # :110  flags.bits()
    out += bytes([depth])               # :111  max_depth
    out += chunk_world_size_be          # :112-114  f32 big-endian: 1.25 = 0x3FA00000
    out += be32(0) + be32(0)            # :115-116  RESERVED_1, RESERVED_2
    out += be32(bounds[0]) + be32(bounds[1]) + be32(bounds[2])   # :118-121
    out += bytes([len(name)]) + name    # :123-124
    out += md5_digest                   # :126-131  MD5 of the UNcompressed payload (RFC 1321)
    out += be32(len(payload_))          # :141-143
    out += payload_                     # :144
    return out

"""The oracle's restatement of VoxModel::serialize / deserialize (world/voxmodel.rs:177-408) on the CPU:
round trip, LEB128 known answers for io/varint.rs, and the plain-Python writer."""
import numpy as np

import vtm_ref
from voxelis_b200 import workloads as wl


def test_varint_known_answers():
    # io/varint.rs:5-17 is unsigned LEB128 (the reference has no tests of its own for it)
    assert vtm_ref.varint(0) == b"\x00" and vtm_ref.varint(127) == b"\x7f"
    assert vtm_ref.varint(128) == b"\x80\x01" and vtm_ref.varint(300) == b"\xac\x02"
    assert vtm_ref.varint(0xFFFFFFFF) == b"\xff\xff\xff\xff\x0f"


def test_round_trip_and_python_writer(oracle_api):
    o = oracle_api
    for dtype, vb in ((wl.U8, 1), (wl.I32, 4)):
        m1, v1 = wl.named_workload("sum", 2, 4, dtype)
        m2, v2 = wl.batch_from_function(4, wl.p_random(4), dtype, 3)
        masks, values = np.concatenate([m1, m2]), np.concatenate([v1, v2])
        masks[0] = 0
        c = o.VoxInterner(16 << 20, dtype)
        roots, _ = c.apply_batches_fresh(4, masks, values)
        pos = np.arange(15, dtype=np.int32).reshape(5, 3) - 4
        data = c.model_serialize(pos, roots)
        d = c.download()
        assert data == vtm_ref.payload_from_pools(d["children"], d["values"], d["refs"], vb, pos, roots)
        c2 = o.VoxInterner(16 << 20, dtype)
        pos2, roots2 = c2.model_deserialize(data)
        assert np.array_equal(pos2, pos) and int(roots2[0]) == 0
        for a, b in zip(roots, roots2):
            assert np.array_equal(c.root_to_vec(int(a), 4), c2.root_to_vec(int(b), 4))
        assert c.stats()["alive_nodes"] == c2.stats()["alive_nodes"]
        # a second generation of the file is identical: ids were already dense and ordered
        assert c2.model_serialize(pos2, roots2) == data

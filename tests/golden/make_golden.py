"""Writes tests/golden/self_regression_digests.json: SHA-256 digests of the oracle's outputs for fixed, seeded inputs of the
steps either side of the path (occupancy planes, voxeliser, terrain generator).  The reference itself cannot be run
here (Rust, no cargo), so these are REGRESSION vectors of the pinned oracle, not reference outputs: they freeze today's
answers so that a later change to the oracle or the generators cannot go unnoticed.
Usage: python tests/golden/make_golden.py   (rewrites the json)"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

import meshes
from oracle import oracle as o
from test_oracle_canonical import oracle_build
from test_oracle_occupancy import cell_offsets, chunk_set
from voxelis_b200 import workloads as wl


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()


def compute():
    out = {}
    for dtype, tag in ((wl.U8, "u8"), (wl.I32, "i32")):
        masks, values = chunk_set(5, dtype)
        it, roots, _ = oracle_build(o, 5, masks, values, dtype, budget=256 << 20)
        for lod in (0, 2):
            S = 32 >> lod
            k = min((64 // S) ** 3, 5)
            got = it.occupancy_masks(roots[:k], 5, cell_offsets(S, k, seed=lod), lod=lod)
            out[f"occupancy_{tag}_d5_lod{lod}"] = digest(got["global"], got["active"], got["material_ids"],
                                                         got["material_counts"], got["per_material"])
    # meshes without transcendental functions (numpy's sin / cos may differ in the last bit between CPU generations)
    for name, (verts, faces) in (("random", meshes.random_triangles(40, 2.5, 0.22, 11)),
                                 ("box", meshes.box((0.25, 0.5, 0.125), (1.5, 1.0, 1.75)))):
        mm = verts.min(0)
        fmap = o.face_chunk_map(5, 1.0, mm, verts, faces)
        hs = []
        for pos, flist in fmap.items():
            has, m, v = o.voxelize_chunk(wl.I32, pos, 5, 1.0, mm, faces[flist], verts)
            hs.append(digest(np.array(pos), np.array([has]), m, v))
        out[f"voxelize_{name}_d5_i32"] = digest(np.array([int(x, 16) % (1 << 62) for x in hs], np.int64))
        out[f"voxelize_{name}_chunks"] = len(fmap)
    m, v = wl.terrain_world((3, 2, 3), 5, "surface_and_below", wl.U8, materials=3)
    out["terrain_3x2x3_d5_u8_3mat"] = digest(wl.height_field(96, 96, height=64), m, v)
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "self_regression_digests.json")
    json.dump(compute(), open(path, "w"), indent=1, sort_keys=True)
    print(open(path).read())

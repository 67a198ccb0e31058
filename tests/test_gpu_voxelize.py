"""The voxeliser on the device (vx_voxelize_plan + vx_voxelize_chunks_device) against the oracle's
restatement of Voxelizer::build_face_to_chunk_map / voxelize_chunk (reference voxelis-voxelize/src/lib.rs:113-249) and
triangle_cube_intersection (voxelis-math/src/lib.rs:3-214): the Batch arrays of every planned chunk are byte-identical
(the f64 tests take the same decisions: same order of operations, no fused multiply-add on either side), and the DAG
built from the generated slab equals the oracle's build of its own batches."""
import numpy as np
import pytest
import torch

import meshes
import parity
from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def mesh_cases():
    s_v, s_f = meshes.uv_sphere((1.45, 1.37, 1.52), 1.21)
    r_v, r_f = meshes.random_triangles(60, 2.5, 0.22, 7)
    big_v = np.array([[0.03, 0.11, 0.07], [2.9, 0.4, 0.3], [0.2, 2.7, 2.2]], np.float64)
    b_v, b_f = meshes.box((0.25, 0.5, 0.125), (1.5, 1.0, 1.75))
    return {"sphere": (s_v, s_f), "random": (r_v, r_f), "big": (big_v, np.array([[1, 2, 3]], np.int32)),
            "box_on_voxel_boundaries": (b_v, b_f)}


def voxelize_on_device(g_api, g, depth, cws, verts, faces, dtype):
    mesh_min = verts.min(0)                                          # Obj::aabb.0 (voxelis-voxelize/src/lib.rs:121)
    plan = g_api.voxelize_plan(depth, cws, mesh_min, verts, faces)
    n = len(plan[0])
    B = wl.blocks_per_chunk(depth)
    dev = torch.device("cuda", 0)
    m = torch.full((n, B, 2), 0xAB, dtype=torch.uint8, device=dev)   # junk: the call must zero the slab itself
    v = torch.full((n, B, 8), 77, dtype=torch.uint8 if dtype == wl.U8 else torch.int32, device=dev)
    hp = torch.full((n,), 9, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    g.voxelize_chunks_device(depth, cws, mesh_min, verts, faces, plan, m.data_ptr(), v.data_ptr(), hp.data_ptr())
    return mesh_min, plan, m, v, hp



@pytest.mark.parametrize("dtype", [wl.I32, wl.U8], ids=["i32", "u8"])
@pytest.mark.parametrize("name", ["sphere", "random", "big", "box_on_voxel_boundaries"])
def test_voxelized_batches_equal_oracle(gpu_api, oracle_api, name, dtype):
    verts, faces = mesh_cases()[name]
    for depth, cws in ((5, 1.0), (4, 0.75)):
        g = gpu_api.VoxInterner.with_memory_budget(64 << 20, dtype)
        mesh_min, (positions, pc, pf), m, v, hp = voxelize_on_device(gpu_api, g, depth, cws, verts, faces, dtype)
        want_map = oracle_api.face_chunk_map(depth, cws, mesh_min, verts, faces)
        assert [tuple(p) for p in positions.tolist()] == list(want_map.keys())
        gm, gv, ghp = m.cpu().numpy(), v.cpu().numpy(), hp.cpu().numpy()
        for c, pos in enumerate(want_map):
            flist = pf[pc == c]
            assert flist.tolist() == want_map[pos]
            has, om, ov = oracle_api.voxelize_chunk(dtype, pos, depth, cws, mesh_min, faces[flist], verts)
            assert bool(ghp[c]) == has, (name, depth, pos)
            assert np.array_equal(gm[c], om), (name, depth, pos)
            assert np.array_equal(gv[c], ov), (name, depth, pos)
        assert ghp.any()


def test_voxelized_mesh_builds_like_the_oracle(gpu_api, oracle_api):
    verts, faces = mesh_cases()["sphere"]
    depth, cws, dtype = 5, 1.0, wl.I32                               # the reference voxelises into Batch<i32>
    g = gpu_api.VoxInterner.with_memory_budget(64 << 20, dtype)
    mesh_min, (positions, pc, pf), m, v, hp = voxelize_on_device(gpu_api, g, depth, cws, verts, faces, dtype)
    keep = torch.nonzero(hp).flatten()                               # chunks without patches are dropped (:245-249)
    mk, vk = m[keep].contiguous(), v[keep].contiguous()
    n = len(keep)
    roots = torch.zeros(n, dtype=torch.int64, device=m.device)
    changed = torch.zeros(n, dtype=torch.uint8, device=m.device)
    torch.cuda.synchronize()
    g.apply_batches_device(depth, n, mk.data_ptr(), vk.data_ptr(), roots.data_ptr(), changed.data_ptr())
    g.sync()
    c = oracle_api.VoxInterner(64 << 20, dtype)
    om, ov = [], []
    for ci in keep.cpu().numpy():
        has, a, b = oracle_api.voxelize_chunk(dtype, positions[ci], depth, cws, mesh_min, faces[pf[pc == ci]], verts)
        assert has
        om.append(a)
        ov.append(b)
    flags, fills = parity.flags_from(n)
    croots, cchanged = c.apply_batches_fresh(depth, np.stack(om), np.stack(ov), flags & 1, fills, (flags >> 1) & 1)
    parity.assert_parity(gpu_api, oracle_api, depth, g, roots.cpu().numpy().astype(np.uint64), changed.cpu().numpy(),
                         c, croots, cchanged)

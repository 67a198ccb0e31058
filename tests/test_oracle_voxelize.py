"""Pins the oracle's restatement of voxelis-math (triangle / cube tests of the voxeliser) on the known answers of the
reference's own unit tests, /root/reference/voxelis-math/src/lib.rs:216-814 (ported assertion by assertion), and
checks voxelize_chunk (voxelis-voxelize/src/lib.rs:159-249) on shapes whose answer is known by construction."""
import numpy as np

from voxelis_b200 import workloads as wl

UNIT = ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))
TRI = ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0))
QUAD = ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0))


def test_point_in_or_on_cube_known_answers(oracle_api):
    """voxelis-math/src/lib.rs:219-300."""
    f = oracle_api.point_in_or_on_cube
    inside = [(0.5, 0.5, 0.5), (0.0, 0.5, 0.5), (1.0, 0.5, 0.5), (0.5, 0.0, 0.5), (0.5, 1.0, 0.5), (0.5, 0.5, 0.0),
              (0.5, 0.5, 1.0), (0.0, 0.0, 0.5), (1.0, 1.0, 0.5), (0.5, 0.0, 0.0), (0.5, 1.0, 1.0), (0.0, 0.0, 0.0),
              (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0), (1.0, 0.0, 1.0), (0.0, 1.0, 1.0),
              (1.0, 1.0, 1.0)]
    outside = [(-0.1, 0.5, 0.5), (1.1, 0.5, 0.5), (0.5, -0.1, 0.5), (0.5, 1.1, 0.5), (0.5, 0.5, -0.1), (0.5, 0.5, 1.1),
               (1.0 + 1e-5, 0.5, 0.5), (0.5, 1.0 + 1e-5, 0.5), (0.5, 0.5, 1.0 + 1e-5)]
    assert all(f(p, UNIT) for p in inside)
    assert not any(f(p, UNIT) for p in outside)


def test_point_in_or_on_triangle_known_answers(oracle_api):
    """voxelis-math/src/lib.rs:303-412."""
    f = oracle_api.point_in_or_on_triangle
    yes = [(0.25, 0.25, 0.0), (0.5, 0.0, 0.0), (0.0, 0.5, 0.0), (0.5, 0.5, 0.0), (0.0, 0.0, 0.0), (1.0, 0.0, 0.0),
           (0.0, 1.0, 0.0), (1.0 / 3.0, 1.0 / 3.0, 0.0)]
    no = [(1.0, 1.0, 0.0), (-0.1, 0.5, 0.0), (0.5, -0.1, 0.0), (1.0 + 1e-5, 0.0, 0.0), (0.0, 1.0 + 1e-5, 0.0),
          (-1e-5, -1e-5, 0.0)]
    assert all(f(p, TRI) for p in yes)
    assert not any(f(p, TRI) for p in no)


def test_edge_quad_intersection_known_answers(oracle_api):
    """voxelis-math/src/lib.rs:417-466 (and point_in_quad :470-640 through the intersection point)."""
    f = oracle_api.edge_quad_intersection
    assert not f(((1.5, 1.5, 0.0), (2.0, 1.5, 0.0)), QUAD)
    assert f(((0.5, 0.5, -0.5), (0.5, 0.5, 0.5)), QUAD)
    assert not f(((0.0, 0.0, 1.0), (1.0, 0.0, 1.0)), QUAD)
    assert not f(((1.0 + 1e-4, 0.5, 0.0), (2.0, 0.5, 0.0)), QUAD)
    for x, y, hit in ((0.5, 0.0, True), (1.0, 0.5, True), (0.0, 0.5, True), (0.0, 0.0, True), (1.0, 1.0, True),
                      (-0.5, 0.5, False), (1.5, 0.5, False), (0.5, -0.5, False), (0.5, 1.5, False)):
        assert f(((x, y, -1.0), (x, y, 1.0)), QUAD) == hit, (x, y)        # a vertical edge through (x, y, 0)


def test_point_in_quad_known_answers(oracle_api):
    """voxelis-math/src/lib.rs:469-652."""
    f = oracle_api.point_in_quad
    yes = [(0.5, 0.5, 0.0), (0.5, 0.0, 0.0), (1.0, 0.5, 0.0), (0.5, 1.0, 0.0), (0.0, 0.5, 0.0), (0.0, 0.0, 0.0),
           (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0), (1.0 - 1e-6, 0.5, 0.0)]
    no = [(-0.5, 0.5, 0.0), (1.5, 0.5, 0.0), (0.5, -0.5, 0.0), (0.5, 1.5, 0.0), (-0.5, -0.5, 0.0), (10.0, 10.0, 10.0),
          (1.0 + 1e-4, 0.5, 0.0)]
    assert all(f(p, QUAD) for p in yes)
    assert not any(f(p, QUAD) for p in no)
    bent = ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 1.0), (0.0, 1.0, 0.0))            # :524-545 non-planar quad
    assert f((0.5, 0.5, 0.25), bent)
    assert not f((0.5, 0.5, 1.0), bent)


def test_triangle_cube_intersection_known_answers(oracle_api):
    """voxelis-math/src/lib.rs:657-812."""
    f = oracle_api.triangle_cube_intersection
    yes = [((0.25, 0.25, 0.25), (0.75, 0.25, 0.25), (0.25, 0.75, 0.25)),
           ((-0.5, 0.5, 0.5), (0.5, 0.5, 0.5), (1.5, 0.5, 0.5)),
           ((-0.5, 0.5, 0.5), (0.5, 0.5, 0.5), (0.5, 1.5, 0.5)),
           ((0.5, 0.5, -0.5), (0.5, 0.5, 0.5), (0.5, 1.5, 0.5)),
           ((-0.5, 0.5, 0.5), (1.5, 0.5, 0.5), (0.5, 2.0, 0.5)),
           ((0.5, 0.5, 0.5), (2.0, 2.0, 2.0), (1.5, 1.5, 2.0)),
           ((0.2, 0.2, 0.2), (0.8, 0.2, 0.2), (0.5, 0.8, 0.2)),
           ((0.5, 0.5, 1.0), (0.75, 0.25, 1.0), (0.25, 0.75, 1.0))]
    no = [((1.5, 1.5, 1.5), (2.5, 1.5, 1.5), (1.5, 2.5, 1.5)),
          ((0.5, 1.5, 1.5), (1.5, 1.5, 1.5), (1.0, 2.0, 1.5)),
          ((1.5, 0.5, 0.5), (2.5, 0.5, 0.5), (2.0, 1.5, 0.5)),
          ((-1.5, -1.5, -1.5), (-1.0, -1.5, -1.5), (-1.5, -1.0, -1.5)),
          ((0.0, 0.0, 1.5), (1.0, 0.0, 1.5), (0.0, 1.0, 1.5)),
          ((0.5, 0.5, 2.0), (1.5, 0.5, 2.0), (0.5, 1.5, 2.0))]
    assert all(f(t, UNIT) for t in yes)
    assert not any(f(t, UNIT) for t in no)
    assert not f(TRI, ((0.5, 0.5, 0.5), (1.5, 1.5, 1.5)))                  # :713-721 cube above the triangle


def test_voxelize_chunk_axis_aligned_plate(oracle_api):
    """A horizontal quad (two faces) at Y = 10.5 voxels over x, z in [4.5, 20.5]: exactly the voxels of layer 10 whose
    cell meets it, nothing else (voxelize_chunk, voxelis-voxelize/src/lib.rs:159-249)."""
    depth, cws = 5, 32.0                                                   # voxel size 1
    verts = np.array([[4.5, 10.5, 4.5], [20.5, 10.5, 4.5], [20.5, 10.5, 20.5], [4.5, 10.5, 20.5]])
    faces = np.array([[1, 2, 3], [1, 3, 4]], np.int32)
    has, masks, values = oracle_api.voxelize_chunk(wl.U8, (0, 0, 0), depth, cws, (0.0, 0.0, 0.0), faces, verts)
    assert has
    dense = wl.dense_expected(masks, values)                               # [y][z][x]
    want = np.zeros((32, 32, 32), np.uint8)
    want[10, 4:21, 4:21] = 1
    assert np.array_equal(dense, want)
    m = oracle_api.face_chunk_map(depth, cws, (0.0, 0.0, 0.0), verts, faces)
    assert list(m.keys()) == [(0, 0, 0)] and m[(0, 0, 0)] == [0, 1]
    has2, masks2, _ = oracle_api.voxelize_chunk(wl.U8, (1, 0, 0), depth, cws, (0.0, 0.0, 0.0), faces, verts)
    assert not has2 and not masks2.any()


def test_host_plan_equals_face_chunk_map(oracle_api):
    """vx_voxelize_plan (host logic of the product, no device involved) == build_face_to_chunk_map restated in Python
    (voxelis-voxelize/src/lib.rs:113-156): same chunks in first-seen order, same face list per chunk."""
    import meshes
    import voxelis_b200 as vx
    for verts, faces in (meshes.uv_sphere((1.45, 1.37, 1.52), 1.21), meshes.random_triangles(60, 2.5, 0.22, 7),
                         meshes.box((0.25, 0.5, 0.125), (1.5, 1.0, 1.75))):
        for depth, cws in ((5, 1.0), (4, 0.75), (6, 2.0)):
            mm = verts.min(0)
            pos, pc, pf = vx.voxelize_plan(depth, cws, mm, verts, faces)
            want = oracle_api.face_chunk_map(depth, cws, mm, verts, faces)
            assert [tuple(p) for p in pos.tolist()] == list(want.keys())
            for c, key in enumerate(want):
                assert pf[pc == c].tolist() == want[key]
    # one triangle across hundreds of chunks: the binding's capacity guess fails and the second call fills the lists
    verts = np.array([[0.0, 0.0, 0.0], [9.7, 0.3, 0.2], [0.4, 9.1, 8.8]], np.float64)
    faces = np.array([[1, 2, 3]], np.int32)
    pos, pc, pf = vx.voxelize_plan(4, 0.5, verts.min(0), verts, faces)
    want = oracle_api.face_chunk_map(4, 0.5, verts.min(0), verts, faces)
    assert len(pos) == len(want) > 66 and [tuple(p) for p in pos.tolist()] == list(want.keys()) and not pf.any()
    with __import__("pytest").raises(vx.VoxelisError):
        vx.voxelize_plan(5, 1.0, (0, 0, 0), np.zeros((2, 3)), np.array([[1, 2, 3]], np.int32))   # vertex 3 does not exist

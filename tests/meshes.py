"""Small synthetic triangle meshes for the voxeliser tests (vertices f64 [nv][3], faces int32 [nf][3], 1-based like
the reference's Obj loader, voxelis-voxelize/src/lib.rs:186-188)."""
import numpy as np


def uv_sphere(center, radius, nu=12, nv=8):
    verts = [[center[0], center[1] + radius, center[2]]]
    for j in range(1, nv):
        th = np.pi * j / nv
        for i in range(nu):
            ph = 2 * np.pi * i / nu
            verts.append([center[0] + radius * np.sin(th) * np.cos(ph), center[1] + radius * np.cos(th),
                          center[2] + radius * np.sin(th) * np.sin(ph)])
    verts.append([center[0], center[1] - radius, center[2]])
    faces = []
    ring = lambda j, i: 1 + (j - 1) * nu + (i % nu)
    for i in range(nu):
        faces.append([0, ring(1, i), ring(1, i + 1)])
        faces.append([len(verts) - 1, ring(nv - 1, i + 1), ring(nv - 1, i)])
    for j in range(1, nv - 1):
        for i in range(nu):
            faces.append([ring(j, i), ring(j + 1, i), ring(j + 1, i + 1)])
            faces.append([ring(j, i), ring(j + 1, i + 1), ring(j, i + 1)])
    return np.array(verts, np.float64), np.array(faces, np.int32) + 1


def random_triangles(n, extent, size, seed):
    rng = np.random.default_rng(seed)
    base = rng.uniform(0, extent, (n, 1, 3))
    verts = (base + rng.uniform(-size, size, (n, 3, 3))).reshape(-1, 3)
    faces = np.arange(3 * n, dtype=np.int32).reshape(n, 3) + 1
    return verts.astype(np.float64), faces


def box(lo, hi):
    """Axis-aligned box: every face lies ON voxel boundaries when lo / hi are multiples of the voxel size — the
    epsilon paths of the tests (voxelis-math/src/lib.rs:17-26,44-47)."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1],
                  [x0, y1, z1]], np.float64)
    q = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (0, 3, 7, 4), (1, 2, 6, 5)]
    f = [[a, b, c] for a, b, c, d in q] + [[a, c, d] for a, b, c, d in q]
    return v, np.array(f, np.int32) + 1

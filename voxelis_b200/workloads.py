"""Synthetic batch generators (host side, numpy) — SURVEY.md §8(d).

Every generator returns batches in the reference's ``Batch`` array layout
(voxelis/src/core/batch.rs:39-45,153-157): ``masks[n][B][2]`` = (set_mask, clear_mask) and
``values[n][B][8]``, ``B = 8**(D-1)``, block index ``p`` = Morton(x>>1, y>>1, z>>1), lane
``i = (x&1) | (y&1)<<1 | (z&1)<<2`` (voxelis/src/utils/common.rs:24-55).

Patterns restate the reference's benchmark inputs (voxelis/benches/voxtree_bench.rs:553-563
uniform, :699-710 sum, :777-786 checkerboard, :1265-1278 terrain surface-only) and test
inputs (voxelis/src/spatial/voxtree.rs:1770-2185).  The terrain height field is this
repo's own integer value-noise (fastnoise-lite is not part of the reference tree): batch
generation is not the hot path, and the same arrays feed the oracle and the GPU.
"""
from __future__ import annotations

import numpy as np

U8, I32 = 0, 1
NP_DTYPE = {U8: np.uint8, I32: np.int32}


def blocks_per_chunk(depth: int) -> int:
    return 1 << (3 * max(depth - 1, 0))


def _spread(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint32) & 0x3FF
    v = (v | (v << 16)) & 0x30000FF
    v = (v | (v << 8)) & 0x300F00F
    v = (v | (v << 4)) & 0x30C30C3
    v = (v | (v << 2)) & 0x9249249
    return v


def morton(x, y, z) -> np.ndarray:
    """encode_child_index_path (utils/common.rs:24-55)."""
    return _spread(np.asarray(x)) | (_spread(np.asarray(y)) << 1) | (_spread(np.asarray(z)) << 2)


_COORDS: dict = {}


def lane_coords(depth: int):
    """(x, y, z) of every batch slot in slot order ``p*8+i``; each array has length 8**depth."""
    if depth not in _COORDS:
        n = 1 << depth
        x, y, z = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        m = morton(x.ravel(), y.ravel(), z.ravel())
        order = np.argsort(m, kind="stable")
        _COORDS[depth] = (x.ravel()[order].astype(np.int32), y.ravel()[order].astype(np.int32),
                          z.ravel()[order].astype(np.int32))
    return _COORDS[depth]


def batch_from_function(depth: int, fn, dtype=U8, n_chunks: int = 1, chunk_arg=None):
    """values = fn(x, y, z, c) per slot; returns (masks, values) with set bits where the
    function returned a *set* decision.  ``fn`` returns (value_array, set_bool_array)."""
    B = blocks_per_chunk(depth)
    x, y, z = lane_coords(depth)
    masks = np.zeros((n_chunks, B, 2), np.uint8)
    values = np.zeros((n_chunks, B, 8), NP_DTYPE[dtype])
    bit = (1 << np.arange(8)).astype(np.uint16)
    for c in range(n_chunks):
        v, s = fn(x, y, z, c if chunk_arg is None else chunk_arg[c])
        v = np.asarray(v).astype(NP_DTYPE[dtype]).reshape(B, 8)
        s = np.broadcast_to(np.asarray(s, bool), (B * 8,)).reshape(B, 8)
        setb = s & (v != 0)
        clrb = s & (v == 0)
        masks[c, :, 0] = (setb * bit).sum(1).astype(np.uint8)
        masks[c, :, 1] = (clrb * bit).sum(1).astype(np.uint8)
        values[c] = np.where(s, v, 0)
    return masks, values


def batch_from_dense(vol_xyz: np.ndarray, set_xyz=None, dtype=U8):
    """One chunk from a dense [x][y][z] volume, as if ``batch.set(pos, v)`` had been called for
    every voxel where ``set_xyz`` is true (default: every voxel)."""
    n = vol_xyz.shape[0]
    depth = int(np.log2(n))
    s_all = np.ones_like(vol_xyz, bool) if set_xyz is None else set_xyz

    def fn(x, y, z, _c):
        return vol_xyz[x, y, z], s_all[x, y, z]

    m, v = batch_from_function(depth, fn, dtype)
    return m[0], v[0]


def dense_expected(masks, values, fill=None):
    """Dense [y][z][x] volume a FRESH tree must unfold to (to_vec layout,
    utils/common.rs:229-238): set voxels take their value, the rest the fill or 0."""
    B = masks.shape[0]
    depth = int(round(np.log2(B * 8) / 3))
    n = 1 << depth
    x, y, z = lane_coords(depth)
    bits = ((masks[:, 0][:, None] >> np.arange(8)) & 1).astype(bool).ravel()
    vals = values.reshape(-1)
    eff = np.where(bits, vals, 0 if fill is None else fill).astype(values.dtype)
    out = np.zeros((n, n, n), values.dtype)
    out[y, z, x] = eff
    return out


# ----------------------------------------------------------------------------- patterns
def p_uniform(v=1):
    return lambda x, y, z, c: (np.full(x.shape, v), True)


def p_uniform_half(v=1):
    def fn(x, y, z, c):
        n = int(round(x.size ** (1 / 3)))
        return np.full(x.shape, v), y < n // 2
    return fn


def p_checkerboard_bench(v=1):
    """voxtree_bench.rs:777-786 — only even-parity voxels are set."""
    return lambda x, y, z, c: (np.full(x.shape, v), ((x + y + z) % 2) == 0)


def p_checkerboard_test():
    """voxtree.rs:1486-1494 — every voxel set: 2 on even parity, 1 on odd."""
    return lambda x, y, z, c: (np.where(((x + y + z) % 2) == 0, 2, 1), True)


def p_sum(offset=1):
    """voxtree_bench.rs:699-710 — v = x+y+z+offset."""
    return lambda x, y, z, c: (x + y + z + offset, True)


def p_sum_per_chunk():
    """SURVEY §8(d) cfg 2: v = ((x+y+z+c) mod 255)+1 so chunks differ."""
    return lambda x, y, z, c: (((x + y + z + c) % 255) + 1, True)


def p_sparse(v=1, step=4):
    """voxtree.rs:1782-1790 — every 4th voxel on each axis."""
    return lambda x, y, z, c: (np.full(x.shape, v), (x % step == 0) & (y % step == 0) & (z % step == 0))


def p_gradient():
    """voxtree.rs:1865-1873 — value = x % 256 (x == 0 is a clear)."""
    return lambda x, y, z, c: (x % 256, True)


def p_hollow_cube(v=1):
    """voxtree.rs:1962-1980 — faces of the cube only."""
    def fn(x, y, z, c):
        n = int(round(x.size ** (1 / 3)))
        face = (x == 0) | (x == n - 1) | (y == 0) | (y == n - 1) | (z == 0) | (z == n - 1)
        return np.full(x.shape, v), face
    return fn


def p_diagonal(v=1):
    """voxtree.rs:2056-2059 — x == y == z."""
    return lambda x, y, z, c: (np.full(x.shape, v), (x == y) & (y == z))


_SM_G = np.uint64(0x9E3779B97F4A7C15)
_SM_A = np.uint64(0xBF58476D1CE4E5B9)
_SM_B = np.uint64(0x94D049BB133111EB)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 output per input counter (uint64, wrapping)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + _SM_G)
        z = (z ^ (z >> np.uint64(30))) * _SM_A
        z = (z ^ (z >> np.uint64(27))) * _SM_B
        return z ^ (z >> np.uint64(31))


SEED_BASE = 0x5EED0000


def p_random(k=255, cell=1):
    """v = 1 + splitmix64 mod k per voxel (k=255 full entropy) — or, with
    ``k=4`` and zero allowed, ``splitmix64 mod 4``.  ``cell`` > 1 makes cell³ voxels share
    one value.  Seed = 0x5EED0000 + chunk index."""
    def fn(x, y, z, c):
        n = int(round(x.size ** (1 / 3)))
        lin = ((x // cell).astype(np.uint64) * np.uint64(n) + (y // cell).astype(np.uint64)) * np.uint64(n) \
            + (z // cell).astype(np.uint64)
        with np.errstate(over="ignore"):
            r = splitmix64(lin + (np.uint64(SEED_BASE + c) << np.uint64(32)))
        if k == 4:
            return (r % np.uint64(4)).astype(np.int64), True
        return (1 + (r % np.uint64(k))).astype(np.int64), True
    return fn


# ------------------------------------------------------------------------------ terrain
def height_field(nx: int, nz: int, seed: int = SEED_BASE, height: int = 256, x0: int = 0, z0: int = 0):
    """Integer 4-octave value noise, h in [0, height).  Octave o has lattice period 256>>o voxels
    and weight 8>>o (sum 15); lattice values are the low 16 bits of
    splitmix64(seed<<40 ^ o<<36 ^ ix<<18 ^ iz); interpolation is 16.16 fixed-point smoothstep."""
    X = (np.arange(nx, dtype=np.int64) + x0)[:, None]
    Z = (np.arange(nz, dtype=np.int64) + z0)[None, :]
    acc = np.zeros((nx, nz), np.int64)
    for o in range(4):
        P = 256 >> o
        w = 8 >> o
        ix, fx = X // P, X % P
        iz, fz = Z // P, Z % P

        def lat(a, b):
            key = (np.uint64(seed) << np.uint64(40)) ^ (np.uint64(o) << np.uint64(36)) ^ \
                  ((a.astype(np.uint64) & np.uint64(0x3FFFF)) << np.uint64(18)) ^ \
                  (b.astype(np.uint64) & np.uint64(0x3FFFF))
            return (splitmix64(key) & np.uint64(0xFFFF)).astype(np.int64)

        def smooth(f):
            t = (f * 65536) // P
            return (t * t * (3 * 65536 - 2 * t)) >> 32

        sx, sz = smooth(fx), smooth(fz)
        c00, c10 = lat(ix + 0 * iz, iz + 0 * ix), lat(ix + 1 + 0 * iz, iz + 0 * ix)
        c01, c11 = lat(ix + 0 * iz, iz + 1 + 0 * ix), lat(ix + 1 + 0 * iz, iz + 1 + 0 * ix)
        a = c00 + (((c10 - c00) * sx) >> 16)
        b = c01 + (((c11 - c01) * sx) >> 16)
        acc += w * (a + (((b - a) * sz) >> 16))
    h16 = acc // 15
    return ((h16 * height) >> 16).astype(np.int32)


def terrain_world(grid=(64, 8, 64), depth=5, variant="surface_only", dtype=U8, seed=SEED_BASE,
                  x_chunk_offset=0, materials=1):
    """"Perlin dunes" world (BASELINE.json config 3; README.md:39): ``grid`` = chunks along
    (x, y, z).  Chunk linear index = (cx * gy + cy) * gz + cz.

    surface_only       one voxel per (X, Z) column at Y = h   (benches :1265-1278; shapes.rs:302-304)
    surface_and_below  every voxel with Y <= h                 (shapes.rs:306-309); with
                       ``materials=3``: 1 for the surface voxel and the two below it, 2 for the next two, 3 deeper (:344-350)
    """
    gx, gy, gz = grid
    n = 1 << depth
    B = blocks_per_chunk(depth)
    H = height_field(gx * n, gz * n, seed, height=gy * n, x0=x_chunk_offset * n)
    lx, ly, lz = lane_coords(depth)
    nchunks = gx * gy * gz
    masks = np.zeros((nchunks, B, 2), np.uint8)
    values = np.zeros((nchunks, B, 8), NP_DTYPE[dtype])
    bit = (1 << np.arange(8)).astype(np.uint8)
    for cx in range(gx):
        for cz in range(gz):
            hc = H[cx * n:(cx + 1) * n, cz * n:(cz + 1) * n][lx, lz]
            for cy in range(gy):
                Y = ly + cy * n
                if variant == "surface_only":
                    s = Y == hc
                    if not s.any():
                        continue
                    v = s.astype(NP_DTYPE[dtype])
                else:
                    s = Y <= hc
                    if not s.any():
                        continue
                    if materials == 3:
                        d = hc - Y
                        v = np.where(d <= 2, 1, np.where(d <= 4, 2, 3)) * s      # shapes.rs:344-350
                    else:
                        v = s.astype(NP_DTYPE[dtype])
                c = (cx * gy + cy) * gz + cz
                sb = s.reshape(B, 8)
                masks[c, :, 0] = (sb * bit).sum(1, dtype=np.uint8)
                values[c] = np.asarray(v, NP_DTYPE[dtype]).reshape(B, 8)
    return masks, values


def named_workload(name: str, n_chunks: int, depth: int = 5, dtype=U8):
    """Workloads by name for bench.py / tests."""
    table = {
        "uniform": (p_uniform(1), False),
        "uniform_half": (p_uniform_half(1), False),
        "checkerboard": (p_checkerboard_bench(1), False),
        "sum": (p_sum(1), False),
        "sum_per_chunk": (p_sum_per_chunk(), True),
        "random255": (p_random(255), True),
        "random4": (p_random(4), True),
        "cell4_random255": (p_random(255, cell=4), True),
        "sparse": (p_sparse(), False),
        "hollow": (p_hollow_cube(), False),
        "diagonal": (p_diagonal(), False),
        "gradient": (p_gradient(), False),
    }
    fn, per_chunk = table[name]
    if per_chunk:
        return batch_from_function(depth, fn, dtype, n_chunks)
    m, v = batch_from_function(depth, fn, dtype, 1)
    return np.repeat(m, n_chunks, 0), np.repeat(v, n_chunks, 0)

"""Third, plain-numpy statement of the greedy mesher's occupancy planes (reference voxelis/src/utils/mesh.rs:
418-513): derived from a dense 64^3 material volume instead of the DAG, so it shares no code or traversal order with
the oracle's restatement or the CUDA kernels."""
import numpy as np

PLANE = 64 * 64


def place(dense_chunks, offsets):
    """dense_chunks[i] = [y][z][x] volume of side S; returns the [y][z][x] 64^3 volume with chunk i at offsets[i]."""
    vol = np.zeros((64, 64, 64), dense_chunks[0].dtype)
    for d, (ox, oy, oz) in zip(dense_chunks, offsets):
        s = d.shape[0]
        vol[oy:oy + s, oz:oz + s, ox:ox + s] = d
    return vol


def planes_of(occ):
    """occ[y][z][x] bool -> 3*4096 words: YZ word[y*64+z] bit x, XZ word[z*64+x] bit y, XY word[y*64+x] bit z."""
    bit = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    o = occ.astype(np.uint64)
    yz = (o * bit[None, None, :]).sum(2, dtype=np.uint64)                    # [y][z]
    xz = (o * bit[:, None, None]).sum(0, dtype=np.uint64)                    # [z][x]
    xy = (o * bit[None, :, None]).sum(1, dtype=np.uint64)                    # [y][x]
    return np.concatenate([yz.ravel(), xz.ravel(), xy.ravel()])


def occupancy_from_dense(vol):
    """dict like oracle.VoxInterner.occupancy_masks / the sorted OccupancyData of build() (mesh.rs:263-285)."""
    occ = vol != 0
    glob = planes_of(occ)
    ys, zs, xs = np.nonzero(occ)
    def axis_mask(a):
        m = np.uint64(0)
        for v in np.unique(a):
            m |= np.uint64(1) << np.uint64(v)
        return m
    xm, ym, zm = axis_mask(xs), axis_mask(ys), axis_mask(zs)
    active = np.array([ym, zm, zm, xm, ym, xm], np.uint64)                   # mesh.rs:451-461
    as_usize = vol.astype(np.int64).astype(np.uint64)                        # `*self as usize`
    ids = np.unique(as_usize[occ])
    counts = np.array([np.count_nonzero(as_usize == i) for i in ids], np.uint64)
    pm = np.stack([planes_of(as_usize == i) for i in ids]) if len(ids) else np.zeros((0, 3 * PLANE), np.uint64)
    return {"global": glob, "active": active, "material_ids": ids, "material_counts": counts, "per_material": pm}


def assert_same(a, b, where=""):
    for k in ("global", "active", "material_ids", "material_counts", "per_material"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (where, k)

"""Hand-derived terrain batches (reference pin for SURVEY §8f-4, terrain half).  Source of every rule: the reference,
/root/reference/voxelis/src/utils/shapes.rs:273-357 and core/batch.rs:145-175; nothing here calls this repository's
generators.

generate_terrain_batch (:273-313) / generate_terrain_batch_3_mats (:315-357), for every column (x, z) of one chunk with
surface height local_y:
    surface_only   just_set((x, local_y, z), 1)                                          :302-304 / :338-340
    else           for y in 0..=local_y: just_set((x, y, z), v)                           :306-309 / :341-353
                   v = 1 (single material);  3 materials: 1 if y >= local_y - 2, 2 if y >= local_y - 4, else 3   :344-350
Batch::just_set (core/batch.rs:145-175): block p = Morton(x>>1, y>>1, z>>1) with x in bit 0, y in bit 1, z in bit 2 of
every 3-bit group (utils/common.rs:24-55), lane i = (x&1) | (y&1)<<1 | (z&1)<<2;  masks[p].set |= 1<<i, values[p][i] = v.

The chunk: MaxDepth 3 (8^3 voxels, 64 blocks).  Heights: every column 0, except column (x=1, z=0) = 7 and
column (x=2, z=5) = 4.
"""
import numpy as np

N = 8
HEIGHTS = np.zeros((N, N), np.int32)        # [x][z]
HEIGHTS[1, 0] = 7
HEIGHTS[2, 5] = 4


def morton3(bx, by, bz):                     # utils/common.rs:24-55, written out for 2 bits per axis (8^3 chunk)
    return ((bx & 1) | (by & 1) << 1 | (bz & 1) << 2) | (((bx >> 1) & 1) | ((by >> 1) & 1) << 1 | ((bz >> 1) & 1) << 2) << 3


def just_set(masks, values, x, y, z, v):     # core/batch.rs:145-175
    p = morton3(x >> 1, y >> 1, z >> 1)
    i = (x & 1) | (y & 1) << 1 | (z & 1) << 2
    masks[p, 0] |= 1 << i
    masks[p, 1] &= ~(1 << i) & 0xFF
    values[p, i] = v


def expected(surface_only: bool, materials: int, np_dtype=np.uint8):
    masks = np.zeros((64, 2), np.uint8)
    values = np.zeros((64, 8), np_dtype)
    for z in range(N):                       # shapes.rs:289-290 loop order (irrelevant for the result)
        for x in range(N):
            local_y = int(HEIGHTS[x, z])
            if surface_only:
                just_set(masks, values, x, local_y, z, 1)
            else:
                for y in range(local_y + 1):
                    v = 1 if materials == 1 else (1 if y >= local_y - 2 else 2 if y >= local_y - 4 else 3)
                    just_set(masks, values, x, y, z, v)
    return masks, values


# Literal answers worked out on paper, independent of the helper above:
#  * column (1, 0), height 7, surface voxel (1, 7, 0): block (0, 3, 0) -> low group y-bit (bit 1) + high group y-bit (bit 4)
#    = 2 + 16 = 18; lane = 1 | 1<<1 | 0 = 3.
#  * column (2, 5), height 4, surface voxel (2, 4, 5): block (1, 2, 2) -> low group x-bit = 1, high group y-bit = 16,
#    high group z-bit = 32 -> 49; lane = 0 | 0 | 1<<2 = 4.
#  * column (0, 0), height 0: voxel (0, 0, 0): block 0, lane 0.  Its x-neighbour (1, 0, 0) belongs to the tall column.
#  * 3 materials, column (1, 0): y = 7, 6, 5 -> 1;  y = 4, 3 -> 2;  y = 2, 1, 0 -> 3.
#    voxel (1, 2, 0): block (0, 1, 0) = 2, lane 1 | 0<<1 = 1 -> values[2][1] = 3;  voxel (1, 3, 0): block 2, lane 3 -> 2;
#    voxel (1, 5, 0): block (0, 2, 0) = 16, lane 1 | 1<<1 = 3 -> 1;  voxel (1, 4, 0): block 16, lane 1 -> 2.
LITERAL_SURFACE = [(18, 3, 1), (49, 4, 1), (0, 0, 1)]                      # (block, lane, value) present in surface_only
LITERAL_SURFACE_ABSENT = [(0, 1), (18, 1), (16, 3)]                         # (block, lane) NOT set in surface_only
LITERAL_3MAT = [(2, 1, 3), (2, 3, 2), (16, 3, 1), (16, 1, 2), (18, 3, 1), (18, 1, 1), (0, 1, 3), (0, 3, 3), (0, 0, 1),
                (49, 4, 1)]
# voxel (2, 0, 5) of the h = 4 column: block (1, 0, 2) = 1 + 32 = 33, lane 0 Morton... lane = 0 | 0 | 1<<2 = 4: y = 0 = local_y - 4 -> 2
LITERAL_3MAT += [(33, 4, 2)]

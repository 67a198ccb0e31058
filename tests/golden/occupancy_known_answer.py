"""Hand-derived occupancy planes (reference pin for SURVEY §8f-3, occupancy half).  Every word below is written out from
/root/reference/voxelis/src/utils/mesh.rs: fill_masks_for_region (:418-513), generate_occupancy_masks (:515-596),
OccupancyDataBuilder::build (:263-285).  Nothing here calls this repository's occupancy code.

Layout (:50-67): three 64 x 64-word planes in one array of 3 * 4096 u64:
    YZ plane at offset 0     word[y * 64 + z], bit x        (index_base_x = PLANE_YZ_OFFSET + y*64 + start_z, |= x_mask)
    XZ plane at offset 4096  word[z * 64 + x], bit y        (index_base_y = PLANE_XZ_OFFSET + z*64 + start_x, |= y_mask)
    XY plane at offset 8192  word[y * 64 + x], bit z        (index_base_z = PLANE_XY_OFFSET + y*64 + start_x, |= z_mask)
global_active (:447-457) = [y_mask, z_mask, z_mask, x_mask, y_mask, x_mask].

Case A — ONE leaf of side 4, material 7, at global position (8, 20, 44):
    a MaxDepth-5 chunk (32^3) whose only set voxels are the aligned cube [8,12) x [20,24) x [12,16), all = 7.  Its eight
    2^3 blocks collapse to Leaf(7) (voxtree.rs:826) and the parent of those eight identical leaves collapses again
    (:1050), so the DAG holds one leaf of side 4 at depth 3; the walk (:553-589) reaches it at pos = (8, 20, 12) and calls
    fill_masks_for_region(builder, offset + pos, 4, 7) with the chunk's offset (0, 0, 32).
      run_mask = (1 << 4) - 1 = 0xF;  x_mask = 0xF << 8;  y_mask = 0xF << 20;  z_mask = 0xF << 44          (:437-446)
      for i in 0..4, j in 0..4 (:459-480):  XZ word[(44+i)*64 + 8 + j] |= y_mask
                                            XY word[(20+i)*64 + 8 + j] |= z_mask
                                            YZ word[(20+i)*64 + 44 + j] |= x_mask
      materials[7] += 4*4*4 = 64                                                                            (:429-433)
"""
import numpy as np

YZ, XZ, XY = 0, 4096, 8192


def case_a():
    x_mask, y_mask, z_mask = 0xF << 8, 0xF << 20, 0xF << 44
    glob = np.zeros(3 * 4096, np.uint64)
    for i in range(4):
        for j in range(4):
            glob[XZ + (44 + i) * 64 + 8 + j] = y_mask
            glob[XY + (20 + i) * 64 + 8 + j] = z_mask
            glob[YZ + (20 + i) * 64 + 44 + j] = x_mask
    active = np.array([y_mask, z_mask, z_mask, x_mask, y_mask, x_mask], np.uint64)
    return {"global": glob, "active": active, "material_ids": [7], "material_counts": [64], "per_material": [glob.copy()]}


# a few of those words spelled out as literals (index -> value)
CASE_A_LITERALS = {XZ + 44 * 64 + 8: 0x0000000000F00000, XZ + 47 * 64 + 11: 0x0000000000F00000,
                   XY + 20 * 64 + 8: 0x0000F00000000000, XY + 23 * 64 + 11: 0x0000F00000000000,
                   YZ + 20 * 64 + 44: 0x0000000000000F00, YZ + 23 * 64 + 47: 0x0000000000000F00}
CASE_A_NONZERO_WORDS = 48                       # 16 words per plane


def case_a_volume():
    vol = np.zeros((32, 32, 32), np.int64)      # [x][y][z]
    vol[8:12, 20:24, 12:16] = 7
    return vol, (0, 0, 32)


"""Case B — two materials in one builder, two chunks:
    chunk 0 (MaxDepth 5, offset (0, 0, 0)):  cube [16,24) x [8,16) x [0,8) = 3  -> one leaf of side 8 at (16, 8, 0)
    chunk 1 (MaxDepth 5, offset (32, 0, 0)): block [0,2)^3 = 9                   -> one leaf of side 2 at global (32, 0, 0)
  side 8, material 3:  x_mask = 0xFF << 16, y_mask = 0xFF << 8, z_mask = 0xFF << 0;  materials[3] = 512
  side 2, material 9:  x_mask = 0x3 << 32,  y_mask = 0x3,       z_mask = 0x3;        materials[9] = 8
  build() sorts the materials by id (:267-268): [3, 9]; global = OR of both; global_active = OR of both triples."""


def case_b():
    def region(arr, sx, sy, sz, side):
        run = (1 << side) - 1
        xm, ym, zm = run << sx, run << sy, run << sz
        for i in range(side):
            for j in range(side):
                arr[XZ + (sz + i) * 64 + sx + j] |= np.uint64(ym)
                arr[XY + (sy + i) * 64 + sx + j] |= np.uint64(zm)
                arr[YZ + (sy + i) * 64 + sz + j] |= np.uint64(xm)
        return np.array([ym, zm, zm, xm, ym, xm], np.uint64)
    p3 = np.zeros(3 * 4096, np.uint64)
    p9 = np.zeros(3 * 4096, np.uint64)
    a3 = region(p3, 16, 8, 0, 8)
    a9 = region(p9, 32, 0, 0, 2)
    return {"global": p3 | p9, "active": a3 | a9, "material_ids": [3, 9], "material_counts": [512, 8], "per_material": [p3, p9]}


CASE_B_LITERALS_GLOBAL = {YZ + 8 * 64 + 0: 0x0000000000FF0000,      # y = 8, z = 0: bits x = 16..23 (material 3)
                          YZ + 0 * 64 + 0: 0x0000000300000000,      # y = 0, z = 0: bits x = 32, 33 (material 9)
                          XZ + 0 * 64 + 16: 0x000000000000FF00,     # z = 0, x = 16: bits y = 8..15
                          XZ + 0 * 64 + 32: 0x0000000000000003,     # z = 0, x = 32: bits y = 0, 1
                          XY + 8 * 64 + 16: 0x00000000000000FF,     # y = 8, x = 16: bits z = 0..7
                          XY + 1 * 64 + 33: 0x0000000000000003}     # y = 1, x = 33: bits z = 0, 1
CASE_B_ACTIVE = [0xFF00 | 0x3, 0xFF | 0x3, 0xFF | 0x3, 0xFF0000 | (0x3 << 32), 0xFF00 | 0x3, 0xFF0000 | (0x3 << 32)]


def case_b_volumes():
    v0 = np.zeros((32, 32, 32), np.int64)
    v0[16:24, 8:16, 0:8] = 3
    v1 = np.zeros((32, 32, 32), np.int64)
    v1[0:2, 0:2, 0:2] = 9
    return [(v0, (0, 0, 0)), (v1, (32, 0, 0))]


"""Case C — the whole-volume branch (:487-512): a MaxDepth-6 chunk filled with 5 is ONE leaf of side 64; every word of
every plane, of the material's planes and of global_active becomes u64::MAX; materials[5] = 64^3 = 262 144."""
CASE_C = {"word": 0xFFFFFFFFFFFFFFFF, "material_ids": [5], "material_counts": [262144]}

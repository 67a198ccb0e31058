// occ_host_check.cu — TEST INFRASTRUCTURE (CPU suite only, never part of libvoxelis_b200.so).
// Steps the per-word source of vx_occupancy.cuh (occ_walk_line / occ_word, __host__ __device__) on the host over
// pools downloaded from the CPU oracle, so that the line walk, the plane / axis mapping and the register-vs-spill
// bookkeeping are checked against the oracle before any GPU time is spent.  The kernels themselves are checked on
// the B200 by tests/test_gpu_occupancy.py.
#include <cstring>
#include <map>
#include <vector>

#include "../../voxelis_b200/csrc/vx_occupancy.cuh"

using namespace vx;

// Morton enumeration of a builder's maximal nodes (host side of the test only): drives occ_region_words and
// occ_block_words, the index functions the shared-memory kernel scatters with.
// Item `item` of a builder = 8^min(ld,2) Morton positions of one cell.  For every maximal non-default node the
// item owns: cube(value, x, y, z, ls) — x, y, z in builder voxels, side 2^ls; a node bigger than the item belongs
// to the item that starts it.  A branch at depth ld - 1 is delivered whole: block(v[8], x, y, z) with
// v[dx | dy<<1 | dz<<2] the value of voxel (x+dx, y+dy, z+dz), default where the child is empty.
template <class T, class Cube, class Block>
__host__ __device__ __forceinline__ void occ_walk_item(const u64* __restrict__ children, const T* __restrict__ values,
                                                       const u64* __restrict__ cell, int ld, u32 item, Cube cube,
                                                       Block block) {
    const int gsh = 6 - ld, G = 1 << gsh;
    const int isz_log = 3 * (ld < 2 ? ld : 2);
    const u32 ipc_log = u32(3 * ld - isz_log);                  // items per cell (log2)
    const u32 c = item >> ipc_log, m0 = (item & ((1u << ipc_log) - 1)) << isz_log;
    const u64 root = VX_OCC_LD(&cell[c]);
    if (root == 0) return;
    const u32 ox = (c & (G - 1)) << ld, oz = ((c >> gsh) & (G - 1)) << ld, oy = (c >> (2 * gsh)) << ld;
    u64 path[7];
    path[0] = root;
    u32 m = m0;
    const u32 end = m0 + (1u << isz_log);
    bool first = true;
    while (m < end) {
        int d = first ? 0 : ld - 1 - (VX_OCC_FFS(int(m)) - 1) / 3;  // deepest stored ancestor still containing m
        u64 node = path[d];
        while (node != 0 && !id_is_leaf(node) && d < ld - 1) {
            const int ci = int(m >> (3 * (ld - 1 - d))) & 7;
            node = VX_OCC_LD(&children[size_t(id_index(node)) * 8 + ci]);
            ++d;
            path[d] = node;
        }
        first = false;
        if (node != 0 && !id_is_leaf(node) && d == ld - 1) {    // a block: 8 voxel-level children at once
            const u64* row = &children[size_t(id_index(node)) * 8];
            u64 ch[8];
            T v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ch[i] = VX_OCC_LD(&row[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ch[i] != 0 ? values[id_index(ch[i])] : T(0);
            const u32 start = m & ~7u;
            block(v, ox + occ_compact3(start), oy + occ_compact3(start >> 1), oz + occ_compact3(start >> 2));
            m = start + 8;
            continue;
        }
        const u32 span = 1u << (3 * (ld - d));
        const u32 start = m & ~(span - 1);
        m = start + span;
        if (node == 0 || start < m0) continue;
        const T v = values[id_index(node)];
        if (v == T(0)) continue;
        cube(v, ox + occ_compact3(start), oy + occ_compact3(start >> 1), oz + occ_compact3(start >> 2), u32(ld - d));
    }
}
__host__ __device__ inline u32 occ_items_per_builder(int ld) {
    return u32(1) << (3 * (6 - ld) + 3 * ld - 3 * (ld < 2 ? ld : 2));
}


template <class T>
static int run(const u64* children, const T* values, const u64* cell, int ld, int max_mat, u64* ids, u64* counts,
               u64* global, u64* pm) {
    std::map<u64, u64> mats;
    const int G = 64 >> ld;
    for (int r = 0; r < OCC_PLANE; ++r) {
        const int y = r >> 6, z = r & 63, cy = y >> ld, cz = z >> ld;
        for (int cx = 0; cx < G; ++cx)
            occ_walk_line<T>(children, values, cell[(size_t(cy) * G + cz) * G + cx], ld, 0, 1, 2, y & ((1 << ld) - 1),
                             z & ((1 << ld) - 1),
                             [&](T v, u64 bits) { mats[occ_material_of<T>(v)] += u64(__builtin_popcountll(bits)); });
    }
    if (int(mats.size()) > max_mat) return -1;
    int n = 0;
    for (auto& e : mats) ids[n] = e.first, counts[n] = e.second, ++n;
    for (int w = 0; w < OCC_ALL; ++w)
        global[w] = occ_word<T>(children, values, cell, ld, w, n, pm + w,
                                [&](T v) { return occ_search(ids, n, occ_material_of<T>(v)); });
    return n;
}

// The shared-memory path's index math: occ_walk_item (Morton node walk, ownership, coordinates, whole blocks),
// occ_region_words (cube -> half words of one plane) and occ_block_words (voxel pairs), stepped item by item; big cubes go through the lane-strided
// form the kernel's warps use.  Materials take slots in the order met and leave in id order.
template <class T>
static int run_planes(const u64* children, const T* values, const u64* cell, int ld, int max_mat, u64* ids,
                      u64* counts, u64* global, u64* pm) {
    std::vector<u64> seen;  // slot -> material id
    std::vector<u64> vol;
    std::vector<std::vector<uint32_t>> planes;  // [plane][(1 + slot) * OCC_HALVES]
    for (int plane = 0; plane < 3; ++plane) {
        std::vector<uint32_t> g(OCC_HALVES, 0);
        std::vector<std::vector<uint32_t>> mats;
        auto slot_of = [&](T v) {
            const u64 id = occ_material_of<T>(v);
            size_t slot = 0;
            while (slot < seen.size() && seen[slot] != id) ++slot;
            if (slot == seen.size()) seen.push_back(id), vol.push_back(0);
            if (mats.size() <= slot) mats.resize(slot + 1, std::vector<uint32_t>(OCC_HALVES, 0));
            return slot;
        };
        for (u32 item = 0; item < occ_items_per_builder(ld); ++item)
            occ_walk_item<T>(
                children, values, cell, ld, item,
                [&](T v, u32 x, u32 y, u32 z, u32 ls) {
                    const size_t slot = slot_of(v);
                    if (plane == 0) vol[slot] += u64(1) << (3 * ls);
                    auto orfn = [&](u32 i, u32 bits) { g[i] |= bits, mats[slot][i] |= bits; };
                    if (ls >= 3)
                        for (u32 lane = 0; lane < 32; ++lane) occ_region_words(plane, x, y, z, ls, lane, 32, orfn);
                    else
                        occ_region_words(plane, x, y, z, ls, 0, 1, orfn);
                },
                [&](const T* v, u32 x, u32 y, u32 z) {
                    occ_block_words<T>(plane, x, y, z, v, [&](T val, u32 i, u32 bits) {
                        const size_t slot = slot_of(val);
                        if (plane == 0) vol[slot] += u64(__builtin_popcount(bits));
                        g[i] |= bits, mats[slot][i] |= bits;
                    });
                });
        mats.resize(seen.size(), std::vector<uint32_t>(OCC_HALVES, 0));
        if (int(seen.size()) > max_mat) return -1;
        std::vector<size_t> order(seen.size());
        for (size_t k = 0; k < seen.size(); ++k) {
            size_t rank = 0;
            for (size_t j = 0; j < seen.size(); ++j) rank += seen[j] < seen[k];
            order[rank] = k;
        }
        memcpy(global + plane * OCC_PLANE, g.data(), OCC_PLANE * 8);
        for (size_t r = 0; r < order.size(); ++r) {
            memcpy(pm + r * OCC_ALL + plane * OCC_PLANE, mats[order[r]].data(), OCC_PLANE * 8);
            ids[r] = seen[order[r]], counts[r] = vol[order[r]];
        }
    }
    return int(seen.size());
}

extern "C" int occ_host_check_planes(const u64* children, const void* values, int dtype, const u64* cell, int ld,
                                     int max_mat, u64* ids, u64* counts, u64* global, u64* pm) {
    return dtype == 0 ? run_planes<u8>(children, (const u8*)values, cell, ld, max_mat, ids, counts, global, pm)
                      : run_planes<int32_t>(children, (const int32_t*)values, cell, ld, max_mat, ids, counts, global, pm);
}

extern "C" int occ_host_check(const u64* children, const void* values, int dtype, const u64* cell, int ld, int max_mat,
                              u64* ids, u64* counts, u64* global, u64* pm) {
    return dtype == 0 ? run<u8>(children, (const u8*)values, cell, ld, max_mat, ids, counts, global, pm)
                      : run<int32_t>(children, (const int32_t*)values, cell, ld, max_mat, ids, counts, global, pm);
}

// occ_host_check.cu — TEST INFRASTRUCTURE (CPU suite only, never part of libvoxelis_b200.so).
// Steps the per-word source of vx_occupancy.cuh (occ_walk_line / occ_word, __host__ __device__) on the host over
// pools downloaded from the CPU oracle, so that the line walk, the plane / axis mapping and the register-vs-spill
// bookkeeping are checked against the oracle before any GPU time is spent.  The kernels themselves are checked on
// the B200 by tests/test_gpu_occupancy.py.
#include <map>
#include <vector>

#include "../../voxelis_b200/csrc/vx_occupancy.cuh"

using namespace vx;

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

extern "C" int occ_host_check(const u64* children, const void* values, int dtype, const u64* cell, int ld, int max_mat,
                              u64* ids, u64* counts, u64* global, u64* pm) {
    return dtype == 0 ? run<u8>(children, (const u8*)values, cell, ld, max_mat, ids, counts, global, pm)
                      : run<int32_t>(children, (const int32_t*)values, cell, ld, max_mat, ids, counts, global, pm);
}

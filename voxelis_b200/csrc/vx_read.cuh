// vx_read.cuh — read-back kernels: VoxTree::get and to_vec over the device pools.
//
// Replaces  get_at_depth  utils/common.rs:122-156  (one thread per query, <= D child hops; a Leaf at
//                                                   any depth answers)
//           to_vec        utils/common.rs:158-246  (dense T[N^3], index = y*N*N + z*N + x)
// These run after the build has been synchronised, so plain cached loads are safe.
#pragma once
#include "vx_device.cuh"

namespace vx {

template <class T>
__device__ __forceinline__ bool descend(const u64* __restrict__ children, const T* __restrict__ values, u64 node,
                                        int depth, int x, int y, int z, T* out) {
    int d = 0;
    while (node != 0) {
        if (d >= depth || id_is_leaf(node)) {
            T v = values[id_index(node)];
            *out = v;
            // get_at_depth: a voxel-level node holding the default value reads as None (:135-141)
            return id_is_leaf(node) && d < depth ? true : v != T(0);
        }
        int sh = depth - d - 1;
        int ci = ((x >> sh) & 1) | (((y >> sh) & 1) << 1) | (((z >> sh) & 1) << 2);
        node = __ldg(&children[size_t(id_index(node)) * 8 + ci]);
        ++d;
    }
    *out = T(0);
    return false;
}

template <class T>
__global__ void get_many_kernel(const u64* __restrict__ children, const T* __restrict__ values, u64 root, int depth,
                                size_t n, const int* __restrict__ xyz, u8* found, long long* out) {
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    T v;
    bool f = descend<T>(children, values, root, depth, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &v);
    found[i] = f;
    out[i] = f ? (long long)v : 0;
}

// dense[r][y][z][x] for n roots; grid-stride over n * N^3 voxels, x fastest -> coalesced stores.
template <class T>
__global__ void to_vec_kernel(const u64* __restrict__ children, const T* __restrict__ values,
                              const u64* __restrict__ roots, size_t n, int depth, T* dense) {
    const size_t N = size_t(1) << depth, vol = N * N * N, total = n * vol;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        size_t r = i / vol, o = i - r * vol;
        int x = int(o & (N - 1)), z = int((o >> depth) & (N - 1)), y = int(o >> (2 * depth));
        T v;
        bool f = descend<T>(children, values, roots[r], depth, x, y, z, &v);
        dense[i] = f ? v : T(0);
    }
}

}  // namespace vx

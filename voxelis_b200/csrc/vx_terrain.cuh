// vx_terrain.cuh — batch generation in device memory: the step BEFORE the build path.
//
// Mirrors   generate_terrain_batch / generate_terrain_batch_3_mats   utils/shapes.rs:273-357
//           (one height per (x, z) column; surface-only or surface-and-below; 1 or 3 materials)
// for a whole grid of chunks at once, writing the reference's Batch arrays (core/batch.rs:39-45,153-157:
// masks[B][2] = (set, clear), values[B][8], block p = Morton(x>>1, y>>1, z>>1), lane = (x&1)|(y&1)<<1|(z&1)<<2)
// straight into HBM, so a world never crosses PCIe as 40 KB per chunk.
//
// The reference samples fastnoise-lite 1.1.1 OpenSimplex2 (float, third-party, not under /root/reference:
// "terrain inputs: parity unpinned by design").  The height field here is this repo's own integer
// 4-octave value noise, bit-identical to voxelis_b200/workloads.py:height_field / terrain_world (numpy), which is
// what tests/test_gpu_terrain.py compares against:
//   octave o: lattice period P = 256 >> o voxels, weight 8 >> o (sum 15); lattice value = low 16 bits of
//   splitmix64(seed << 40 ^ o << 36 ^ (ix & 0x3FFFF) << 18 ^ (iz & 0x3FFFF)); 16.16 fixed-point smoothstep.
//
//   terrain_heights_kernel   one thread per column, heights[x][z] (int32), coalesced along z
//   terrain_batches_kernel   one thread per block of 2x2x2 voxels: 4 heights in, 2 mask bytes + 8 values out;
//                            consecutive threads write consecutive blocks -> fully coalesced stores
#pragma once
#include "vx_device.cuh"

namespace vx {

__host__ __device__ inline u64 splitmix64(u64 x) {
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ long long terrain_lattice(u64 seed, int o, long long a, long long b) {
    const u64 key = (seed << 40) ^ (u64(o) << 36) ^ ((u64(a) & 0x3FFFFull) << 18) ^ (u64(b) & 0x3FFFFull);
    return (long long)(splitmix64(key) & 0xFFFFull);
}

__device__ __forceinline__ long long terrain_smooth(long long f, long long P) {
    const long long t = (f * 65536) / P;
    return (t * t * (3 * 65536 - 2 * t)) >> 32;
}

__global__ void terrain_heights_kernel(u32 nx, u32 nz, u64 seed, u32 height, long long x0, long long z0,
                                       int* __restrict__ heights) {
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= size_t(nx) * nz) return;
    const long long X = (long long)(i / nz) + x0, Z = (long long)(i % nz) + z0;
    long long acc = 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const long long P = 256 >> o, w = 8 >> o;
        const long long ix = X / P, fx = X % P, iz = Z / P, fz = Z % P;
        const long long sx = terrain_smooth(fx, P), sz = terrain_smooth(fz, P);
        const long long c00 = terrain_lattice(seed, o, ix, iz), c10 = terrain_lattice(seed, o, ix + 1, iz);
        const long long c01 = terrain_lattice(seed, o, ix, iz + 1), c11 = terrain_lattice(seed, o, ix + 1, iz + 1);
        const long long a = c00 + (((c10 - c00) * sx) >> 16);
        const long long b = c01 + (((c11 - c01) * sx) >> 16);
        acc += w * (a + (((b - a) * sz) >> 16));
    }
    const long long h16 = acc / 15;
    heights[i] = int((h16 * (long long)height) >> 16);
}

__device__ __forceinline__ u32 compact10(u32 v) {  // inverse of spread10 (utils/common.rs:24-55)
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030C30C3u;
    v = (v | (v >> 4)) & 0x0300F00Fu;
    v = (v | (v >> 8)) & 0x030000FFu;
    v = (v | (v >> 16)) & 0x000003FFu;
    return v;
}

// chunk linear index = (cx * gy + cy) * gz + cz; heights[(cx*n + x) * (gz*n) + cz*n + z]
template <class T>
__global__ void terrain_batches_kernel(int depth, u32 gx, u32 gy, u32 gz, const int* __restrict__ heights,
                                       int surface_only, int materials, u8* __restrict__ masks,
                                       T* __restrict__ values) {
    const int bshift = 3 * (depth - 1);
    const size_t total = (size_t(gx) * gy * gz) << bshift;
    const u32 n = 1u << depth, hz = gz * n;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t c = i >> bshift;
        const u32 p = u32(i & ((size_t(1) << bshift) - 1));
        const u32 cz = u32(c % gz), cy = u32((c / gz) % gy), cx = u32(c / (size_t(gz) * gy));
        const u32 x = compact10(p) * 2, y = compact10(p >> 1) * 2, z = compact10(p >> 2) * 2;
        const int* hp = heights + size_t(cx * n + x) * hz + cz * n + z;
        const int h00 = hp[0], h01 = hp[1], h10 = hp[hz], h11 = hp[hz + 1];  // [dx][dz]
        const int Y0 = int(cy * n + y);
        u32 set = 0;
        T v[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const int h = (l & 1) ? ((l & 4) ? h11 : h10) : ((l & 4) ? h01 : h00);
            const int d = h - (Y0 + ((l >> 1) & 1));
            const bool s = surface_only ? d == 0 : d >= 0;
            const int val = !s ? 0 : (surface_only || materials != 3) ? 1 : d == 0 ? 1 : d <= 3 ? 2 : 3;
            set |= u32(s) << l;
            v[l] = T(val);
        }
        reinterpret_cast<uchar2*>(masks)[i] = make_uchar2((unsigned char)set, 0);
        if (sizeof(T) == 1) {
            u64 pack = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) pack |= u64(u8(v[l])) << (8 * l);
            reinterpret_cast<u64*>(values)[i] = pack;
        } else {
            int4* dst = reinterpret_cast<int4*>(values + i * 8);
            dst[0] = make_int4(int(v[0]), int(v[1]), int(v[2]), int(v[3]));
            dst[1] = make_int4(int(v[4]), int(v[5]), int(v[6]), int(v[7]));
        }
    }
}

}  // namespace vx

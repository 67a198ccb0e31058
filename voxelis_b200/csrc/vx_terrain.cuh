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
//   terrain_batches_kernel   one thread per 4 Morton-consecutive blocks (u8) / per block (i32): column heights in,
//                            mask bytes + values out; consecutive threads write consecutive bytes
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

__device__ __forceinline__ void st_v8(void* p, uint4 a, uint4 b) {  // one 32-byte store (sm_100+)
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// value of a voxel standing d = h - Y below the surface of its column (shapes.rs:302-309,340-355)
__device__ __forceinline__ int terrain_value(int d, bool surface_only, bool three) {
    if (surface_only) return d == 0;
    if (d < 0) return 0;
    return !three ? 1 : d <= 2 ? 1 : d <= 4 ? 2 : 3;  // shapes.rs:344-350: y >= h-2 -> 1, y >= h-4 -> 2, else 3
}

// chunk linear index = (cx * gy + cy) * gz + cz; heights[(cx*n + x) * (gz*n) + cz*n + z].
// One thread = QUAD Morton-consecutive blocks (QUAD = 4: the 2x2 blocks (bx, by) in {0,1}^2 of one bz, i.e. a
// 4 x 4 x 2 voxel tile; QUAD = 1: one block): its 2*QUAD column heights are read once, its 8*QUAD values leave as
// one 32-byte store (u8, QUAD 4) / two 16-byte stores (i32, QUAD 1).  Tiles wholly above the surface or wholly
// in the deep material skip the per-voxel evaluation.
template <class T, int QUAD>
__global__ void __launch_bounds__(256)
terrain_batches_kernel(int depth, u32 gx, u32 gy, u32 gz, const int* __restrict__ heights, int surface_only,
                       int materials, u8* __restrict__ masks, T* __restrict__ values) {
    constexpr int NX = QUAD == 4 ? 4 : 2, NY = QUAD == 4 ? 2 : 1;  // columns along x, block rows along y
    const int bshift = 3 * (depth - 1);
    const size_t total = ((size_t(gx) * gy * gz) << bshift) / QUAD;
    const size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const size_t i = t * QUAD;                                   // first block of the tile (global block index)
    const size_t c = i >> bshift;
    const u32 p = u32(i & ((size_t(1) << bshift) - 1));
    const u32 n = 1u << depth, hz = gz * n;
    const u32 cz = u32(c % gz), cy = u32((c / gz) % gy), cx = u32(c / (size_t(gz) * gy));
    const u32 x = compact10(p) * 2, y = compact10(p >> 1) * 2, z = compact10(p >> 2) * 2;
    const int2* hp = reinterpret_cast<const int2*>(heights + size_t(cx * n + x) * hz + cz * n + z);  // z is even
    int h[NX][2], hmin = 0x7FFFFFFF, hmax = -0x7FFFFFFF - 1;
#pragma unroll
    for (int k = 0; k < NX; ++k) {
        const int2 v = __ldg(hp + size_t(k) * (hz / 2));
        h[k][0] = v.x, h[k][1] = v.y;
        hmin = min(hmin, min(v.x, v.y)), hmax = max(hmax, max(v.x, v.y));
    }
    const int Y0 = int(cy * n + y), Y1 = Y0 + 2 * NY - 1;        // the tile spans Y0 .. Y1
    const bool three = materials == 3, so = surface_only != 0;
    u32 set[QUAD];
    u32 val[QUAD][2];  // 8 lanes x 8 bits per block (values are 0..3)
    if (Y0 > hmax || (so && Y1 < hmin)) {                         // nothing of this tile is set
#pragma unroll
        for (int q = 0; q < QUAD; ++q) set[q] = 0, val[q][0] = val[q][1] = 0;
    } else if (!so && hmin - Y1 >= (three ? 5 : 0)) {             // every voxel is the deep material (d >= 5)
        const u32 f = three ? 0x03030303u : 0x01010101u;
#pragma unroll
        for (int q = 0; q < QUAD; ++q) set[q] = 0xFF, val[q][0] = val[q][1] = f;
    } else {
#pragma unroll
        for (int q = 0; q < QUAD; ++q) {                          // block q: bx = q & 1, by = q >> 1
            u32 sb = 0, lo = 0, hi = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const int d = h[2 * (q & 1) + (l & 1)][l >> 2] - (Y0 + 2 * (q >> 1) + ((l >> 1) & 1));
                const u32 vv = u32(terrain_value(d, so, three));
                sb |= u32(vv != 0) << l;
                if (l < 4) lo |= vv << (8 * l); else hi |= vv << (8 * (l - 4));
            }
            set[q] = sb, val[q][0] = lo, val[q][1] = hi;
        }
    }
    if (QUAD == 4) {
        *reinterpret_cast<uint2*>(masks + i * 2) = make_uint2(set[0] | (set[1] << 16), set[2] | (set[3] << 16));
    } else {
        reinterpret_cast<uchar2*>(masks)[i] = make_uchar2((unsigned char)set[0], 0);
    }
    if (sizeof(T) == 1) {
        if (QUAD == 4) {
            st_v8(values + i * 8, make_uint4(val[0][0], val[0][1], val[1][0], val[1][1]),
                  make_uint4(val[2][0], val[2][1], val[3][0], val[3][1]));
        } else {
            *reinterpret_cast<uint2*>(values + i * 8) = make_uint2(val[0][0], val[0][1]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < QUAD; ++q) {
            int4* dst = reinterpret_cast<int4*>(values + (i + q) * 8);
            dst[0] = make_int4(val[q][0] & 255, (val[q][0] >> 8) & 255, (val[q][0] >> 16) & 255, val[q][0] >> 24);
            dst[1] = make_int4(val[q][1] & 255, (val[q][1] >> 8) & 255, (val[q][1] >> 16) & 255, val[q][1] >> 24);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// High-entropy benchmark batches in device memory: bit-identical to
// voxelis_b200/workloads.py:p_random —  v = 1 + splitmix64(lin + (seed_base + chunk) << 32) mod k  per cell of
// `cell`^3 voxels (lin = ((x/cell)*N + y/cell)*N + z/cell), or, with k == 4, splitmix64 mod 4 with 0 = "set to the
// default value" (clear bit, batch.rs:162-168).  One thread per block of 2^3 voxels: 2 mask bytes + 8 values.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void random_batches_kernel(u32 depth, u64 total_blocks, u32 blocks_log, u64 seed_base, u64 chunk0, u32 k, u32 cell,
                                      u8* __restrict__ masks, T* __restrict__ values) {
    const u64 t = blockIdx.x * u64(blockDim.x) + threadIdx.x;
    if (t >= total_blocks) return;
    const u64 chunk = t >> blocks_log;
    const u32 p = u32(t & ((1ull << blocks_log) - 1));
    const u32 bx = compact10(p), by = compact10(p >> 1), bz = compact10(p >> 2);
    const u64 N = 1ull << depth;
    const u64 key = (seed_base + chunk0 + chunk) << 32;
    u32 setm = 0, clrm = 0;
    T v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const u64 x = 2 * bx + (i & 1), y = 2 * by + ((i >> 1) & 1), z = 2 * bz + (i >> 2);
        const u64 lin = ((x / cell) * N + (y / cell)) * N + (z / cell);
        const u64 r = splitmix64(lin + key);
        const u32 val = k == 4 ? u32(r % 4) : 1u + u32(r % k);
        v[i] = T(val);
        setm |= u32(val != 0) << i;
        clrm |= u32(val == 0) << i;
    }
    *reinterpret_cast<uchar2*>(masks + t * 2) = make_uchar2(u8(setm), u8(clrm));
    if (sizeof(T) == 1) {
        u64 w = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) w |= u64(u8(v[i])) << (8 * i);
        *reinterpret_cast<u64*>(values + t * 8) = w;
    } else {
        uint4* o = reinterpret_cast<uint4*>(values + t * 8);
        o[0] = make_uint4(u32(v[0]), u32(v[1]), u32(v[2]), u32(v[3]));
        o[1] = make_uint4(u32(v[4]), u32(v[5]), u32(v[6]), u32(v[7]));
    }
}

}  // namespace vx

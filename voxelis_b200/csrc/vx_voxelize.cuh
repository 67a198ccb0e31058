// vx_voxelize.cuh — triangle -> voxel batches in device memory (second half).
//
// Replaces  Voxelizer::voxelize_chunk     voxelis-voxelize/src/lib.rs:159-249  (per chunk: every face of the chunk's
//                                          list against every voxel of the face's clipped bounding box)
//           triangle_cube_intersection    voxelis-math/src/lib.rs:3-127   and its helpers :129-214
// writing the reference's Batch arrays (core/batch.rs:39-45,145-175: set bit + value 1) of ALL planned chunks straight
// into the slab vx_apply_batches_device reads.
//
// Everything is f64 in the reference's order of operations (glam 0.29.3 vector helpers: dot = x*x' + y*y' + z*z',
// cross = (y*z' - y'*z, z*x' - z'*x, x*y' - x'*y), normalize = v * (1 / sqrt(dot))); the library is built with
// -fmad=false so that no multiply-add is fused, double division and sqrt are IEEE by default: every decision is the
// one the CPU restatement takes.  The triangle's normal and plane offset do not depend on the voxel and are computed
// once per (chunk, face) pair.
//
//   voxelize_pairs_kernel   one warp per (chunk, face) pair: the 32 lanes stride over the voxels of the clipped box;
//                           a hit ORs the voxel's bit into the block's set_mask (32-bit atomicOr on the word holding
//                           the byte — several faces and lanes meet in one block) and stores the value 1.
#pragma once
#include "vx_device.cuh"

namespace vx {

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ D3 min3(D3 a, D3 b) { return {a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
__device__ __forceinline__ D3 max3(D3 a, D3 b) { return {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
__device__ __forceinline__ double len3(D3 a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ double signum64(double v) { return v != v ? v : (signbit(v) ? -1.0 : 1.0); }  // f64::signum

__device__ __forceinline__ bool pt_in_or_on_cube(D3 p, D3 cmin, D3 cmax) {  // voxelis-math lib.rs:129-151
    const double size = len3(cmax - cmin), eps = size * 1e-8;
    if (size < 1e-8) return len3(p - cmin) < eps;
    return p.x >= cmin.x - eps && p.x <= cmax.x + eps && p.y >= cmin.y - eps && p.y <= cmax.y + eps &&
           p.z >= cmin.z - eps && p.z <= cmax.z + eps;
}

__device__ __forceinline__ bool pt_in_or_on_triangle(D3 p, D3 a, D3 b, D3 c) {  // :153-178
    const D3 v0 = b - a, v1 = c - a, v2 = p - a;
    const double d00 = dot3(v0, v0), d01 = dot3(v0, v1), d02 = dot3(v0, v2), d11 = dot3(v1, v1), d12 = dot3(v1, v2);
    const double denom = d00 * d11 - d01 * d01;
    if (fabs(denom) < 1e-8) return false;
    const double inv = 1.0 / denom;
    const double u = (d11 * d02 - d01 * d12) * inv;
    const double v = (d00 * d12 - d01 * d02) * inv;
    return u >= 0.0 && v >= 0.0 && (u + v) <= 1.0;
}

__device__ __forceinline__ bool edge_quad(D3 e1, D3 e2, D3 q0, D3 q1, D3 q2, D3 q3) {  // :180-214
    const D3 c = cross3(q1 - q0, q2 - q0);
    const D3 normal = c * (1.0 / len3(c));
    const double denom = dot3(normal, e2 - e1);
    if (fabs(denom) < 1e-8) return false;
    const double t = dot3(normal, q0 - e1) / denom;
    if (!(t >= 0.0 && t <= 1.0)) return false;
    const D3 p = e1 + t * (e2 - e1);
    return pt_in_or_on_triangle(p, q0, q1, q2) || pt_in_or_on_triangle(p, q0, q2, q3);
}

// triangle_cube_intersection — voxelis-math lib.rs:3-127; tri_min / tri_max / normal / d hoisted by the caller
__device__ __noinline__ bool tri_cube(D3 tv0, D3 tv1, D3 tv2, D3 tri_min, D3 tri_max, D3 normal, double d, D3 cmin,
                                      D3 cmax) {
    const double eps = 1e-5;
    if (tri_max.x < cmin.x - eps || tri_min.x > cmax.x + eps || tri_max.y < cmin.y - eps || tri_min.y > cmax.y + eps ||
        tri_max.z < cmin.z - eps || tri_min.z > cmax.z + eps)
        return false;
    D3 cp[8] = {d3(cmin.x, cmin.y, cmin.z), d3(cmax.x, cmin.y, cmin.z), d3(cmax.x, cmax.y, cmin.z),
                d3(cmin.x, cmax.y, cmin.z), d3(cmin.x, cmin.y, cmax.z), d3(cmax.x, cmin.y, cmax.z),
                d3(cmax.x, cmax.y, cmax.z), d3(cmin.x, cmax.y, cmax.z)};
    const double sign = signum64(dot3(normal, cp[0]) + d);
    for (int i = 1; i < 8; ++i) {
        const double s = dot3(normal, cp[i]) + d;
        if (fabs(s) < eps) continue;
        if (signum64(s) != sign) return true;
    }
    if (pt_in_or_on_cube(tv0, cmin, cmax) || pt_in_or_on_cube(tv1, cmin, cmax) || pt_in_or_on_cube(tv2, cmin, cmax))
        return true;
    for (int i = 0; i < 8; ++i)
        if (pt_in_or_on_triangle(cp[i], tv0, tv1, tv2)) return true;
    const int fq[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {0, 3, 7, 4}, {1, 2, 6, 5}};
    for (int e = 0; e < 3; ++e) {
        const D3 e1 = e == 0 ? tv0 : e == 1 ? tv1 : tv2, e2 = e == 0 ? tv1 : e == 1 ? tv2 : tv0;
        for (int f = 0; f < 6; ++f)
            if (edge_quad(e1, e2, cp[fq[f][0]], cp[fq[f][1]], cp[fq[f][2]], cp[fq[f][3]])) return true;
    }
    return false;
}

__device__ __forceinline__ u32 spread10_dev(u32 v) {  // utils/common.rs:24-55
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x30000FF;
    v = (v | (v << 8)) & 0x300F00F;
    v = (v | (v << 4)) & 0x30C30C3;
    v = (v | (v << 2)) & 0x9249249;
    return v;
}

__device__ __forceinline__ int clamp_voxel(double v, int vpa) {  // `as i32` (saturating) then clamp(0, vpa - 1)
    const int i = v != v ? 0 : v >= 2147483647.0 ? 2147483647 : v <= -2147483648.0 ? (-2147483647 - 1) : int(v);
    return i < 0 ? 0 : i > vpa - 1 ? vpa - 1 : i;
}

template <class T>
__global__ void __launch_bounds__(256)
voxelize_pairs_kernel(int depth, double chunk_world_size, double mmx, double mmy, double mmz,
                      const double* __restrict__ vertices, const int* __restrict__ faces,
                      const int* __restrict__ positions, const u32* __restrict__ pair_chunk,
                      const u32* __restrict__ pair_face, size_t n_pairs, u8* __restrict__ masks, T* __restrict__ values,
                      u8* __restrict__ has_patches) {
    const int lane = threadIdx.x & 31;
    const size_t warp0 = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5, nwarps = (size_t(gridDim.x) * blockDim.x) >> 5;
    const int vpa = 1 << depth;
    const size_t B = size_t(1) << (3 * (depth - 1));
    const double voxel_size = chunk_world_size / double(vpa);  // voxelize_mesh, lib.rs:262
    const double epsilon = voxel_size * 1e-7;
    const D3 splat = d3(epsilon, epsilon, epsilon), mesh_min = d3(mmx, mmy, mmz);
    for (size_t pr = warp0; pr < n_pairs; pr += nwarps) {
        const u32 c = pair_chunk[pr], f = pair_face[pr];
        const int* fi = faces + 3 * size_t(f);
        auto vert = [&](int i) { const double* p = vertices + 3 * size_t(i - 1); return d3(p[0], p[1], p[2]); };
        const D3 v1 = vert(fi[0]) - mesh_min, v2 = vert(fi[1]) - mesh_min, v3 = vert(fi[2]) - mesh_min;
        const D3 cw_min = d3(double(positions[3 * c]), double(positions[3 * c + 1]), double(positions[3 * c + 2])) * chunk_world_size;
        const D3 cw_max = cw_min + d3(chunk_world_size, chunk_world_size, chunk_world_size);
        const D3 face_min = min3(min3(v1, v2), v3), face_max = max3(max3(v1, v2), v3);
        const D3 omin = max3(face_min, cw_min) - splat, omax = min3(face_max, cw_max) + splat;
        if (omin.x >= omax.x || omin.y >= omax.y || omin.z >= omax.z) continue;  // lib.rs:200-206
        const D3 lo = (omin - cw_min) / voxel_size, hi = (omax - cw_min) / voxel_size;
        const int x0 = clamp_voxel(floor(lo.x), vpa), y0 = clamp_voxel(floor(lo.y), vpa), z0 = clamp_voxel(floor(lo.z), vpa);
        const int x1 = clamp_voxel(ceil(hi.x), vpa), y1 = clamp_voxel(ceil(hi.y), vpa), z1 = clamp_voxel(ceil(hi.z), vpa);
        if (x1 < x0 || y1 < y0 || z1 < z0) continue;
        const int nx = x1 - x0 + 1, nz = z1 - z0 + 1, total = nx * nz * (y1 - y0 + 1);
        const D3 normal = cross3(v2 - v1, v3 - v1);
        const double d = -dot3(normal, v1);
        bool any = false;
        for (int k = lane; k < total; k += 32) {
            const int x = x0 + k % nx, z = z0 + (k / nx) % nz, y = y0 + k / (nx * nz);
            const D3 wp = cw_min + d3(double(x), double(y), double(z)) * voxel_size;
            const D3 wmin = wp - splat, wmax = wp + d3(voxel_size, voxel_size, voxel_size) + splat;
            if (!tri_cube(v1, v2, v3, face_min, face_max, normal, d, wmin, wmax)) continue;
            const u32 full = spread10_dev(u32(x)) | (spread10_dev(u32(y)) << 1) | (spread10_dev(u32(z)) << 2);
            const size_t blk = size_t(c) * B + (full >> 3);
            const uintptr_t addr = reinterpret_cast<uintptr_t>(masks + blk * 2);   // set_mask byte of the block
            atomicOr(reinterpret_cast<u32*>(addr & ~uintptr_t(3)), (1u << (full & 7)) << (8 * (addr & 3)));
            values[blk * 8 + (full & 7)] = T(1);
            any = true;
        }
        if (any) has_patches[c] = 1;
    }
}

}  // namespace vx

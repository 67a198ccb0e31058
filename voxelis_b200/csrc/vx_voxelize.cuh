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
//   voxelize_pairs_kernel   one warp per (chunk, face) pair for the cheap first part of the test (bounding boxes,
//                           plane straddle): the 32 lanes stride over the voxels of the clipped box; undecided voxels
//                           are parked per warp across pairs and the long part (vertex in cube, corner in triangle,
//                           18 edge / face tests) always runs on a full warp.  A hit ORs the voxel's bit into the
//                           block's set_mask (32-bit atomicOr on the word holding the byte — several faces and lanes
//                           meet in one block) and stores the value 1.
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

// triangle_cube_intersection — voxelis-math lib.rs:3-127, cut in two at :52.  tri_min / tri_max / normal / d do not
// depend on the cube and are hoisted by the caller.
//   tri_cube_quick  :9-50   bounding boxes apart -> 0 (no); the triangle's plane separates two cube corners -> 1 (yes);
//                           otherwise 2: undecided, the long part has to run
//   tri_cube_slow   :52-126 a triangle vertex in the cube, a cube corner in the triangle, an edge through a cube face
__device__ __forceinline__ int tri_cube_quick(D3 tri_min, D3 tri_max, D3 normal, double d, D3 cmin, D3 cmax) {
    const double eps = 1e-5;
    if (tri_max.x < cmin.x - eps || tri_min.x > cmax.x + eps || tri_max.y < cmin.y - eps || tri_min.y > cmax.y + eps ||
        tri_max.z < cmin.z - eps || tri_min.z > cmax.z + eps)
        return 0;
    const double sign = signum64(dot3(normal, d3(cmin.x, cmin.y, cmin.z)) + d);
#pragma unroll
    for (int i = 1; i < 8; ++i) {  // corners in the reference's order: (x: i in {1,2,5,6}, y: {2,3,6,7}, z: i >= 4) max
        const D3 p = d3((i == 1 || i == 2 || i == 5 || i == 6) ? cmax.x : cmin.x,
                        (i == 2 || i == 3 || i == 6 || i == 7) ? cmax.y : cmin.y, i >= 4 ? cmax.z : cmin.z);
        const double s = dot3(normal, p) + d;
        if (fabs(s) < eps) continue;
        if (signum64(s) != sign) return 1;
    }
    return 2;
}

__device__ __noinline__ bool tri_cube_slow(D3 tv0, D3 tv1, D3 tv2, D3 cmin, D3 cmax) {
    if (pt_in_or_on_cube(tv0, cmin, cmax) || pt_in_or_on_cube(tv1, cmin, cmax) || pt_in_or_on_cube(tv2, cmin, cmax))
        return true;
    D3 cp[8] = {d3(cmin.x, cmin.y, cmin.z), d3(cmax.x, cmin.y, cmin.z), d3(cmax.x, cmax.y, cmin.z),
                d3(cmin.x, cmax.y, cmin.z), d3(cmin.x, cmin.y, cmax.z), d3(cmax.x, cmin.y, cmax.z),
                d3(cmax.x, cmax.y, cmax.z), d3(cmin.x, cmax.y, cmax.z)};
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

// Everything about one (chunk, face) pair that does not depend on the voxel.
struct PairCtx {
    D3 v1, v2, v3, fmin, fmax, normal, cw_min;
    double d;
    int x0, y0, z0, nx, nz, total;
    u32 c;
};

struct VoxelizeArgs {
    int depth;
    double chunk_world_size, mmx, mmy, mmz;
    const double* vertices;
    const int* faces;
    const int* positions;
    const u32* pair_chunk;
    const u32* pair_face;
    size_t n_pairs;
};

// voxelize_chunk, lib.rs:184-221: the face's bounding box clipped to the chunk -> the voxel range to test
__device__ __forceinline__ bool pair_context(const VoxelizeArgs& a, size_t pr, PairCtx& P) {
    const int vpa = 1 << a.depth;
    const double voxel_size = a.chunk_world_size / double(vpa);  // voxelize_mesh, lib.rs:262
    const double epsilon = voxel_size * 1e-7;
    const D3 splat = d3(epsilon, epsilon, epsilon), mesh_min = d3(a.mmx, a.mmy, a.mmz);
    P.c = a.pair_chunk[pr];
    const int* fi = a.faces + 3 * size_t(a.pair_face[pr]);
    auto vert = [&](int i) { const double* p = a.vertices + 3 * size_t(i - 1); return d3(p[0], p[1], p[2]); };
    P.v1 = vert(fi[0]) - mesh_min, P.v2 = vert(fi[1]) - mesh_min, P.v3 = vert(fi[2]) - mesh_min;
    P.cw_min = d3(double(a.positions[3 * P.c]), double(a.positions[3 * P.c + 1]), double(a.positions[3 * P.c + 2])) * a.chunk_world_size;
    const D3 cw_max = P.cw_min + d3(a.chunk_world_size, a.chunk_world_size, a.chunk_world_size);
    P.fmin = min3(min3(P.v1, P.v2), P.v3), P.fmax = max3(max3(P.v1, P.v2), P.v3);
    const D3 omin = max3(P.fmin, P.cw_min) - splat, omax = min3(P.fmax, cw_max) + splat;
    if (omin.x >= omax.x || omin.y >= omax.y || omin.z >= omax.z) return false;  // lib.rs:200-206
    const D3 lo = (omin - P.cw_min) / voxel_size, hi = (omax - P.cw_min) / voxel_size;
    P.x0 = clamp_voxel(floor(lo.x), vpa), P.y0 = clamp_voxel(floor(lo.y), vpa), P.z0 = clamp_voxel(floor(lo.z), vpa);
    const int x1 = clamp_voxel(ceil(hi.x), vpa), y1 = clamp_voxel(ceil(hi.y), vpa), z1 = clamp_voxel(ceil(hi.z), vpa);
    if (x1 < P.x0 || y1 < P.y0 || z1 < P.z0) return false;
    P.nx = x1 - P.x0 + 1, P.nz = z1 - P.z0 + 1, P.total = P.nx * P.nz * (y1 - P.y0 + 1);
    P.normal = cross3(P.v2 - P.v1, P.v3 - P.v1);
    P.d = -dot3(P.normal, P.v1);
    return true;
}

// voxel k of the pair's box (x fastest, then z, then y — the reference's loop nest :224-226) and its cube :227-233
__device__ __forceinline__ void pair_voxel(const VoxelizeArgs& a, const PairCtx& P, int k, int* x, int* y, int* z, D3* wmin,
                                           D3* wmax) {
    const double voxel_size = a.chunk_world_size / double(1 << a.depth), epsilon = voxel_size * 1e-7;
    const D3 splat = d3(epsilon, epsilon, epsilon);
    *x = P.x0 + k % P.nx, *z = P.z0 + (k / P.nx) % P.nz, *y = P.y0 + k / (P.nx * P.nz);
    const D3 wp = P.cw_min + d3(double(*x), double(*y), double(*z)) * voxel_size;
    *wmin = wp - splat;
    *wmax = wp + d3(voxel_size, voxel_size, voxel_size) + splat;
}

template <class T>
__device__ __forceinline__ void voxel_hit(int depth, u32 c, int x, int y, int z, u8* masks, T* values, u8* has_patches) {
    const size_t B = size_t(1) << (3 * (depth - 1));
    const u32 full = spread10_dev(u32(x)) | (spread10_dev(u32(y)) << 1) | (spread10_dev(u32(z)) << 2);
    const size_t blk = size_t(c) * B + (full >> 3);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(masks + blk * 2);   // set_mask byte of the block
    atomicOr(reinterpret_cast<u32*>(addr & ~uintptr_t(3)), (1u << (full & 7)) << (8 * (addr & 3)));
    values[blk * 8 + (full & 7)] = T(1);  // batch.just_set(pos, 1), lib.rs:239
    has_patches[c] = 1;
}

// One warp per (chunk, face) pair for the quick part of the test (lanes stride over the voxels of the clipped box);
// voxels the quick part leaves undecided are parked in a per-warp queue ACROSS pairs and the long part runs on 32 of
// them at a time, every lane rebuilding its own pair's context — so the expensive code always has a full warp,
// however small the boxes are.
// 250 registers, one CTA per SM: capping at 128 (two CTAs) spills 472 bytes per thread and measured 8 % slower.
template <class T>
__global__ void __launch_bounds__(256)
voxelize_pairs_kernel(VoxelizeArgs a, u8* __restrict__ masks, T* __restrict__ values, u8* __restrict__ has_patches) {
    __shared__ u32 q_pair[8][64], q_k[8][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t warp0 = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5, nwarps = (size_t(gridDim.x) * blockDim.x) >> 5;
    int qn = 0;  // entries parked by this warp (same value in every lane)
    auto slow_round = [&](int count) {  // the last `count` (<= 32) parked voxels
        if (lane < count) {
            PairCtx Q;
            pair_context(a, q_pair[w][qn - count + lane], Q);
            int x, y, z;
            D3 wmin, wmax;
            pair_voxel(a, Q, int(q_k[w][qn - count + lane]), &x, &y, &z, &wmin, &wmax);
            if (tri_cube_slow(Q.v1, Q.v2, Q.v3, wmin, wmax)) voxel_hit<T>(a.depth, Q.c, x, y, z, masks, values, has_patches);
        }
        qn -= count;
        __syncwarp();
    };
    for (size_t pr = warp0; pr < a.n_pairs; pr += nwarps) {
        PairCtx P;
        if (!pair_context(a, pr, P)) continue;
        for (int base = 0; base < P.total; base += 32) {
            const int k = base + lane;
            int st = 0;
            if (k < P.total) {
                int x, y, z;
                D3 wmin, wmax;
                pair_voxel(a, P, k, &x, &y, &z, &wmin, &wmax);
                st = tri_cube_quick(P.fmin, P.fmax, P.normal, P.d, wmin, wmax);
                if (st == 1) voxel_hit<T>(a.depth, P.c, x, y, z, masks, values, has_patches);
            }
            const unsigned pend = __ballot_sync(FULL, st == 2);
            if (st == 2) {
                const int slot = qn + __popc(pend & ((1u << lane) - 1));
                q_pair[w][slot] = u32(pr);
                q_k[w][slot] = u32(k);
            }
            qn += __popc(pend);
            __syncwarp();
            if (qn >= 32) slow_round(32);
        }
    }
    if (qn > 0) slow_round(qn);
}

}  // namespace vx

// vx_occupancy.cuh — the greedy mesher's input, straight from the device pools.
//
// Replaces  generate_occupancy_masks   utils/mesh.rs:515-596  (stack walk of one chunk's DAG)
//           fill_masks_for_region      utils/mesh.rs:418-513  (OR a cube into 3 bit planes, global + per material)
//           OccupancyDataBuilder::build utils/mesh.rs:263-285 (materials sorted by id, masks in that order)
//
// One OccupancyDataBuilder covers a 64^3 voxel volume (mesh.rs:50-67): three bit planes of 64x64 words,
//   YZ  word[y*64 + z] bit x        XZ  4096 + word[z*64 + x] bit y        XY  8192 + word[y*64 + x] bit z
// filled from up to (64/S)^3 chunks of side S = 2^depth placed at multiples of S.  The reference scatters:
// every leaf region ORs S'^2 words in each plane.  Here the work is turned around so that no word is ever
// touched by two threads: one thread OWNS one word of one builder and walks the DAG along the word's bit axis,
// run by run (a node standing at depth k answers 2^(depth-k) bits at once, so collapsed subtrees cost one
// step).  Results are OR / sum reductions, hence independent of the visiting order -> bit-identical to the
// reference's scatter.  HBM traffic = the output planes written once (no memset, no atomics on them) + the
// DAG nodes on the way (L1/L2 hits: neighbouring words share their paths).
//
// These two kernels serve builders with MANY materials (> 5); the usual case goes through occ_planes_kernel
// (shared-memory planes, Morton node walk) at the end of this file.
//
//   occ_materials_kernel   one CTA per builder: voxel count per material (mesh.rs:428-432) from the YZ rows,
//                          ids sorted ascending as build() does; also clears the builder's global_active
//   occ_masks_kernel       48 CTAs per builder x 256 threads = 12288 words: global plane word, per-material
//                          words (written once each, for every material of the builder), global_active by a
//                          CTA-wide OR + one atomicOr per CTA
#pragma once
#include "vx_device.cuh"

namespace vx {

constexpr int OCC_AXIS = 64;
constexpr int OCC_PLANE = OCC_AXIS * OCC_AXIS;
constexpr int OCC_ALL = 3 * OCC_PLANE;
constexpr u32 OCC_ERR_MATERIALS = 1;  // more materials in a builder than the caller made room for

// material_id = `*self as usize` (core/voxel.rs:85-87): sign-extending for signed T
template <class T>
__host__ __device__ __forceinline__ u64 occ_material_of(T v) {
    return u64((long long)v);
}

// Walk one line of 2^ld voxels of the tree `root` along axis `caxis` (0 = x, 1 = y, 2 = z); p / q are the
// local coordinates on the two other axes, (pax, qax) their bit positions in a child index
// (child = x | y<<1 | z<<2, utils/common.rs:104-111).  emit(value, bits) gets each non-default run, bits
// relative to the start of the line.  A branch standing at depth ld answers with its LOD value, like
// `node_id.is_branch() && depth < max_depth` in mesh.rs:562,580.
#ifdef __CUDA_ARCH__
#define VX_OCC_LD(p) __ldg(p)
#define VX_OCC_FFS(c) __ffs(c)
#else  // host instantiation: tests/cpp/occ_host_check.cu steps the same per-word code without a GPU
#define VX_OCC_LD(p) (*(p))
#define VX_OCC_FFS(c) __builtin_ffs(c)
#endif

template <class T, class F>
__host__ __device__ __forceinline__ void occ_walk_line(const u64* __restrict__ children, const T* __restrict__ values,
                                              u64 root, int ld, int caxis, int pax, int qax, int p, int q, F emit) {
    if (root == 0) return;
    u64 path[7];
    path[0] = root;
    const int S = 1 << ld;
    int c = 0;
    while (c < S) {
        // deepest stored node that still contains voxel c: the previous run ended on a 2^ctz(c) boundary
        int d = c ? ld - VX_OCC_FFS(c) : 0;
        u64 node = path[d];
        while (node != 0 && !id_is_leaf(node) && d < ld) {
            const int sh = ld - 1 - d;
            const int ci = (((c >> sh) & 1) << caxis) | (((p >> sh) & 1) << pax) | (((q >> sh) & 1) << qax);
            node = VX_OCC_LD(&children[size_t(id_index(node)) * 8 + ci]);
            ++d;
            path[d] = node;
        }
        const int len = 1 << (ld - d);
        if (node != 0) {
            const T v = values[id_index(node)];
            if (v != T(0)) emit(v, (len == 64 ? ~0ull : ((1ull << len) - 1)) << c);
        }
        c += len;
    }
}

template <class T>
struct OccTable {
    static constexpr int SIZE = sizeof(T) == 1 ? 256 : 2048;  // u8: direct-mapped by value; wider T: hashed
};

// cells[b][(cy*G + cz)*G + cx] = root of the chunk placed at cell (cx, cy, cz) of builder b, 0 = none.
template <class T>
__global__ void __launch_bounds__(256)
occ_materials_kernel(const u64* __restrict__ children, const T* __restrict__ values, const u64* __restrict__ cells,
                     int ld, u32 max_materials, u32* __restrict__ n_materials, u64* __restrict__ material_ids,
                     u64* __restrict__ material_counts, u64* __restrict__ active, u32* __restrict__ err,
                     const u32* __restrict__ only) {
    constexpr int TS = OccTable<T>::SIZE;
    if (only && !only[blockIdx.x]) return;  // this builder went through occ_planes_kernel
    __shared__ u32 s_key[TS];   // raw value bits, 0 = free (the default value is never a material)
    __shared__ u32 s_cnt[TS];
    __shared__ u32 s_over;
    const int b = blockIdx.x, G = OCC_AXIS >> ld, gsh = 6 - ld;
    const u64* cell = cells + (size_t(b) << (3 * gsh));
    for (int i = threadIdx.x; i < TS; i += blockDim.x) {
        s_key[i] = 0;
        s_cnt[i] = 0;
    }
    if (threadIdx.x == 0) s_over = 0;
    if (threadIdx.x < 6) active[size_t(b) * 6 + threadIdx.x] = 0;
    __syncthreads();

    auto add = [&](T v, u64 bits) {
        const u32 n = u32(__popcll(bits));
        if (sizeof(T) == 1) {
            const u32 k = u32(v) & 0xFFu;
            s_key[k] = k;
            atomicAdd(&s_cnt[k], n);
        } else {
            const u32 k = u32(v);
            u32 h = u32(mix64(k)) & (TS - 1);
            for (int probe = 0; probe < TS; ++probe) {
                u32 cur = s_key[h];
                if (cur == 0) {
                    const u32 prev = atomicCAS(&s_key[h], 0u, k);
                    cur = prev == 0 ? k : prev;
                }
                if (cur == k) {
                    atomicAdd(&s_cnt[h], n);
                    return;
                }
                h = (h + 1) & (TS - 1);
            }
            s_over = 1;
        }
    };
    // YZ rows: y = r >> 6, z = r & 63, bits along x — every voxel of the volume is seen exactly once
    for (int r = threadIdx.x; r < OCC_PLANE; r += blockDim.x) {
        const int y = r >> 6, z = r & 63;
        const int cy = y >> ld, cz = z >> ld, ly = y & ((1 << ld) - 1), lz = z & ((1 << ld) - 1);
        for (int cx = 0; cx < G; ++cx) {
            const u64 root = __ldg(&cell[(size_t(cy) * G + cz) * G + cx]);
            occ_walk_line<T>(children, values, root, ld, 0, 1, 2, ly, lz, add);
        }
    }
    __syncthreads();
    // build(): sort by material id (as usize).  Rank of an entry = number of smaller ids present.
    u32 mine = 0;
    for (int i = threadIdx.x; i < TS; i += blockDim.x) mine += s_key[i] != 0;
    __shared__ u32 s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    if (mine) atomicAdd(&s_total, mine);
    __syncthreads();
    const u32 total = s_total;
    if (threadIdx.x == 0) {
        n_materials[b] = total;
        if (total > max_materials || s_over) atomicExch(err, OCC_ERR_MATERIALS);
    }
    if (total > max_materials || s_over) return;
    for (int i = threadIdx.x; i < TS; i += blockDim.x) {
        const u32 k = s_key[i];
        if (k == 0) continue;
        const u64 id = occ_material_of<T>(T(k));
        u32 rank = 0;
        for (int j = 0; j < TS; ++j) {
            const u32 o = s_key[j];
            rank += (o != 0 && occ_material_of<T>(T(o)) < id);
        }
        material_ids[size_t(b) * max_materials + rank] = id;
        material_counts[size_t(b) * max_materials + rank] = s_cnt[i];
    }
}

// One word of one builder: walks the word's line through every cell along the bit axis, returns the global
// plane word and writes the word's column of the per-material planes (pm + slot * OCC_ALL), each exactly once
// when the line holds <= K materials.  slot_of(value) = index in the builder's sorted material list.
template <class T, class SlotOf>
__host__ __device__ __forceinline__ u64 occ_word(const u64* __restrict__ children, const T* __restrict__ values,
                                                 const u64* __restrict__ cell, int ld, int w, int nmat, u64* pm,
                                                 SlotOf slot_of) {
    constexpr int K = 4;  // materials of one word kept in registers; more spill to read-modify-write
    const int G = OCC_AXIS >> ld, S = 1 << ld;
    const int plane = w >> 12, a = (w >> 6) & 63, bb = w & 63;
    // plane 0 (YZ): a = y, bb = z, bits x | plane 1 (XZ): a = z, bb = x, bits y | plane 2 (XY): a = y, bb = x, bits z
    const int caxis = plane == 0 ? 0 : plane == 1 ? 1 : 2;
    const int pax = plane == 0 ? 1 : plane == 1 ? 2 : 1;   // axis of `a`
    const int qax = plane == 0 ? 2 : 0;                    // axis of `bb`
    const int ca = a >> ld, cb = bb >> ld, la = a & (S - 1), lb = bb & (S - 1);

    u64 gmask = 0;
    int ns = 0, sl[K];
    u64 mk[K];
    bool spilled = false;
    int last_slot = -1;
    T last_v = T(0);
#pragma unroll
    for (int i = 0; i < K; ++i) sl[i] = -1, mk[i] = 0;

    for (int cc = 0; cc < G; ++cc) {
        int cx, cy, cz;
        if (plane == 0) cx = cc, cy = ca, cz = cb;
        else if (plane == 1) cx = cb, cy = cc, cz = ca;
        else cx = cb, cy = ca, cz = cc;
        const u64 root = VX_OCC_LD(&cell[(size_t(cy) * G + cz) * G + cx]);
        const int base = cc << ld;
        occ_walk_line<T>(children, values, root, ld, caxis, pax, qax, la, lb, [&](T v, u64 bits) {
            bits <<= base;
            gmask |= bits;
            if (v != last_v) {
                last_v = v;
                last_slot = slot_of(v);
            }
            const int slot = last_slot;
            if (slot < 0 || slot >= nmat) return;  // only after OCC_ERR_MATERIALS
            if (spilled) {
                pm[size_t(slot) * OCC_ALL] |= bits;
                return;
            }
            bool done = false;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (!done && sl[i] == slot) mk[i] |= bits, done = true;
            if (done) return;
            if (ns < K) {
#pragma unroll
                for (int i = 0; i < K; ++i)
                    if (i == ns) sl[i] = slot, mk[i] = bits;
                ++ns;
                return;
            }
            // a fifth material on this word: the word's column goes to memory and is OR-ed in place from now on
            for (int s = 0; s < nmat; ++s) pm[size_t(s) * OCC_ALL] = 0;
#pragma unroll
            for (int i = 0; i < K; ++i) pm[size_t(sl[i]) * OCC_ALL] = mk[i];
            pm[size_t(slot) * OCC_ALL] = bits;
            spilled = true;
        });
    }
    if (!spilled)
        for (int s = 0; s < nmat; ++s) {
            u64 m = 0;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (sl[i] == s) m = mk[i];
            pm[size_t(s) * OCC_ALL] = m;
        }
    return gmask;
}

// slot of a material id in a sorted id list
__host__ __device__ __forceinline__ int occ_search(const u64* __restrict__ ids, int n, u64 id) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (VX_OCC_LD(&ids[mid]) < id) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <class T>
__global__ void __launch_bounds__(256)
occ_masks_kernel(const u64* __restrict__ children, const T* __restrict__ values, const u64* __restrict__ cells,
                 int ld, u32 max_materials, const u32* __restrict__ n_materials,
                 const u64* __restrict__ material_ids, u64* __restrict__ global, u64* __restrict__ active,
                 u64* __restrict__ per_material, const u32* __restrict__ only) {
    if (only && !only[blockIdx.y]) return;
    __shared__ u8 s_lut[256];
    __shared__ u64 s_or[8];
    const int b = blockIdx.y, gsh = 6 - ld;
    const int nmat = int(min(n_materials[b], max_materials));
    const u64* ids = material_ids + size_t(b) * max_materials;
    if (sizeof(T) == 1) {
        for (int i = threadIdx.x; i < nmat; i += blockDim.x) s_lut[u32(ids[i]) & 0xFFu] = u8(i);
        __syncthreads();
    }
    const u64* cell = cells + (size_t(b) << (3 * gsh));
    const int w = blockIdx.x * blockDim.x + threadIdx.x;  // word of the builder, 0 .. 12287; one plane per CTA
    const int plane = w >> 12;
    u64* pm = per_material + (size_t(b) * max_materials) * OCC_ALL + w;  // + slot * OCC_ALL
    const u64 gmask = occ_word<T>(children, values, cell, ld, w, nmat, pm, [&](T v) -> int {
        if (sizeof(T) == 1) return int(s_lut[u32(v) & 0xFFu]);
        return occ_search(ids, nmat, occ_material_of<T>(v));
    });
    global[size_t(b) * OCC_ALL + w] = gmask;
    // global_active (mesh.rs:451-461): [0] y, [1] z, [2] z, [3] x, [4] y, [5] x = OR of the run masks, i.e. the OR of
    // every word whose bits run along that axis
    u64 o = gmask;
#pragma unroll
    for (int k = 16; k; k >>= 1) o |= __shfl_xor_sync(FULL, o, k);
    if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = o;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int i = 0; i < int(blockDim.x >> 5); ++i) t |= s_or[i];
        if (t) {
            const int i0 = plane == 0 ? 3 : plane == 1 ? 0 : 1;
            const int i1 = plane == 0 ? 5 : plane == 1 ? 4 : 2;
            atomicOr((ull*)&active[size_t(b) * 6 + i0], (ull)t);
            atomicOr((ull*)&active[size_t(b) * 6 + i1], (ull)t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory path: builders with at most `ms` (<= 5) materials — the mesher's normal case.
//
// One CTA per (builder, plane) keeps the builder's per-material planes of that axis in shared memory (32 KiB per
// material; the global plane is their OR and is formed while streaming out).  The DAG is unfolded level by level,
// node by node instead of row by row, with the frontier in shared memory: 8 lanes read the 64-byte child row of a
// node (one child each), uniform cubes of side >= 8 are queued and later filled by whole warps (an aligned run
// of <= 32 bits lives in one 32-bit half word), branches go to the next frontier.  Below the 8^3 nodes 64 lanes
// take one node, every lane one 2^3 block (4^3 child -> block): its 8 child ids and their 8 values are independent
// loads, and the two voxels that share a
// half word leave as one shared atomic when they agree — all lanes run the same code, whatever the tree looks
// like.  Needs depth - lod >= 3.  Materials take slots in the order the CTA meets them; the planes
// leave in id order (build(), mesh.rs:263-285) through a permutation applied while they are streamed out with
// 16-byte coalesced stores — each output byte is written exactly once, nothing is read back from HBM.
// The YZ CTA also delivers the material list and voxel counts; every CTA stores its two global_active words.
// A builder with more than `ms` materials is flagged in overflow[] and left to the word-owner kernels above.

__host__ __device__ __forceinline__ u32 occ_compact3(u32 v) {  // every third bit of a Morton index (<= 10 bits out)
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030C30C3u;
    v = (v | (v >> 4)) & 0x0300F00Fu;
    v = (v | (v >> 8)) & 0x030000FFu;
    v = (v | (v >> 16)) & 0x000003FFu;
    return v;
}

// OR the cube (x, y, z, side 2^ls) into one plane held as 32-bit half words: half word 2*w + (bit >> 5) of word w.
// or32(index, bits) does the OR (a shared atomic on the device, a plain OR in the host stepping test).
template <class F>
__host__ __device__ __forceinline__ void occ_region_words(int plane, u32 x, u32 y, u32 z, u32 ls, u32 k0,
                                                          u32 kstep, F or32) {
    const u32 side = 1u << ls;
    const u32 r0 = plane == 1 ? z : y, r1 = plane == 0 ? z : x, rb = plane == 0 ? x : plane == 1 ? y : z;
    const u32 bits = side >= 32 ? 0xFFFFFFFFu : (((1u << side) - 1) << (rb & 31));
    const u32 half = (rb >> 5) & 1;
    for (u32 k = k0; k < side * side; k += kstep) {
        const u32 w = (r0 + (k >> ls)) * 64 + r1 + (k & (side - 1));
        if (side == 64) {
            or32(2 * w, bits);
            or32(2 * w + 1, bits);
        } else {
            or32(2 * w + half, bits);
        }
    }
}

// The 8 voxels of a block -> 4 half words of one plane; the two voxels along the bit axis share a half word and
// leave together when they hold the same material.  orv(value, index, bits).
template <class T, class F>
__host__ __device__ __forceinline__ void occ_block_words(int plane, u32 x, u32 y, u32 z, const T* v, F orv) {
    const u32 rb = plane == 0 ? x : plane == 1 ? y : z;         // even: the pair sits at bits rb, rb + 1
    const u32 half = (rb >> 5) & 1, sh = rb & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if ((i >> plane) & 1) continue;                         // i = the voxel of the pair at bit rb
        const int j = i | (1 << plane);
        const u32 dx = i & 1, dy = (i >> 1) & 1, dz = i >> 2;
        const u32 r0 = plane == 1 ? z + dz : y + dy, r1 = plane == 0 ? z + dz : x + dx;
        const u32 idx = 2 * (r0 * 64 + r1) + half;
        if (v[i] != T(0) && v[i] == v[j]) {
            orv(v[i], idx, 3u << sh);
        } else {
            if (v[i] != T(0)) orv(v[i], idx, 1u << sh);
            if (v[j] != T(0)) orv(v[j], idx, 2u << sh);
        }
    }
}

constexpr int OCC_MS_MAX = 5;        // materials per builder on the shared-memory path
constexpr int OCC_QUEUE = 512;       // cubes of side >= 8 in one 64^3 volume
constexpr int OCC_HALVES = 2 * OCC_PLANE;  // u32 half words per plane

template <class T>
__global__ void __launch_bounds__(1024)
occ_planes_kernel(const u64* __restrict__ children, const T* __restrict__ values, const u64* __restrict__ cells,
                  int ld, int ms, u32 max_materials, u32* __restrict__ n_materials, u64* __restrict__ material_ids,
                  u64* __restrict__ material_counts, u64* __restrict__ global, u64* __restrict__ active,
                  u64* __restrict__ per_material, u32* __restrict__ overflow, u32* __restrict__ overflow_count) {
    extern __shared__ uint4 occ_smem[];
    u32* planes = reinterpret_cast<u32*>(occ_smem);  // [ms][OCC_HALVES]: one plane per material slot
    __shared__ u32 u_idx[2][512];                    // frontier, ping-pong: branch nodes of side >= 8 (<= 512 per level)
    __shared__ u16 u_pos[2][512];                    // x4 | y4 << 4 | z4 << 8: the node's origin in units of 4 voxels
    __shared__ u32 s_mat[OCC_MS_MAX];                // raw value bits of the material in each slot, 0 = free
    __shared__ u32 s_cnt[OCC_MS_MAX];
    __shared__ u32 s_queue[OCC_QUEUE];
    __shared__ u32 s_qn, s_over, s_nu[2], s_act[2];
    __shared__ u8 s_lut[256];                        // u8 values: slot + 1 of a material already met, 0 = not yet
    __shared__ int s_order[OCC_MS_MAX], s_n;
    const int plane = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const int gsh = 6 - ld, G = 1 << gsh;
    const u64* cell = cells + (size_t(b) << (3 * gsh));
    for (int i = tid; i < ms * (OCC_HALVES / 4); i += nthr) occ_smem[i] = make_uint4(0, 0, 0, 0);
    if (tid < OCC_MS_MAX) s_mat[tid] = 0, s_cnt[tid] = 0;
    if (tid < 256) s_lut[tid] = 0;
    if (tid == 0) s_qn = 0, s_over = 0, s_nu[0] = 0, s_nu[1] = 0, s_act[0] = 0, s_act[1] = 0;
    __syncthreads();

    T last_v = T(0);
    int last_slot = -1;
    u32 cnt[OCC_MS_MAX];                                         // voxels per slot seen by this thread (YZ CTA only)
#pragma unroll
    for (int k = 0; k < OCC_MS_MAX; ++k) cnt[k] = 0;
    auto count = [&](int slot, u32 n) {
#pragma unroll
        for (int k = 0; k < OCC_MS_MAX; ++k)
            if (k == slot) cnt[k] += n;
    };
    auto slot_of = [&](T v) -> int {                             // slot of the material, claimed on first sight
        if (sizeof(T) == 1) {
            const int s = s_lut[u32(v) & 0xFFu];
            if (s) return s - 1;
        } else if (v == last_v) {
            return last_slot;
        }
        const u32 key = sizeof(T) == 1 ? (u32(v) & 0xFFu) : u32(v);
        int slot = -1;
        for (int k = 0; k < ms && slot < 0; ++k) {
            u32 cur = s_mat[k];
            if (cur == 0) {
                const u32 prev = atomicCAS(&s_mat[k], 0u, key);
                cur = prev == 0 ? key : prev;
            }
            if (cur == key) slot = k;
        }
        if (slot < 0) s_over = 1;
        if (sizeof(T) == 1 && slot >= 0) s_lut[key] = u8(slot + 1);  // every writer stores the same byte
        last_v = v;
        last_slot = slot;
        return slot;
    };
    // a uniform cube of side 2^ls >= 8 at (x4, y4, z4) * 4: filled by whole warps after the walk
    auto big_cube = [&](u64 id, u32 pos, u32 ls) {
        const T v = values[id_index(id)];
        if (v == T(0)) return;
        const int slot = slot_of(v);
        if (slot < 0) return;
        if (plane == 0) count(slot, 1u << (3 * ls));
        const u32 q = atomicAdd(&s_qn, 1u);
        if (q < OCC_QUEUE)
            s_queue[q] = ((pos & 15) << 2) | (((pos >> 4) & 15) << 8) | (((pos >> 8) & 15) << 14) | (ls << 18) |
                         (u32(slot) << 21);
    };
    // 1. the cells' roots (depth 0, side 2^ld >= 8)
    for (int c = tid; c < G * G * G; c += nthr) {
        const u64 root = __ldg(&cell[c]);
        if (root == 0) continue;
        const u32 un = u32(ld - 2);                              // log2 of the cell side in units of 4 voxels
        const u32 pos = ((c & (G - 1)) << un) | ((u32(c >> (2 * gsh)) << un) << 4) | ((((c >> gsh) & (G - 1)) << un) << 8);
        if (id_is_leaf(root)) {
            big_cube(root, pos, u32(ld));
        } else {
            const u32 q = atomicAdd(&s_nu[0], 1u);
            u_idx[0][q] = id_index(root);
            u_pos[0][q] = u16(pos);
        }
    }
    __syncthreads();
    // 2. level by level down to the 8^3 nodes: 8 lanes per node, one child each (one 64-byte row per node)
    for (int k = 0; k <= ld - 4; ++k) {
        const int src = k & 1, dst = src ^ 1;
        const u32 nsrc = s_nu[src];
        const u32 cls = u32(ld - k - 1);                         // log2 of the child side, >= 3
        for (u32 t = tid; t < nsrc * 8; t += nthr) {
            const u32 e = t >> 3, ci = t & 7;
            const u64 id = __ldg(&children[size_t(u_idx[src][e]) * 8 + ci]);
            if (id == 0) continue;
            const u32 un = cls - 2;
            const u32 pos = u32(u_pos[src][e]) + (((ci & 1) << un) | ((((ci >> 1) & 1) << un) << 4) | (((ci >> 2) << un) << 8));
            if (id_is_leaf(id)) {
                big_cube(id, pos, cls);
            } else {
                const u32 q = atomicAdd(&s_nu[dst], 1u);
                u_idx[dst][q] = id_index(id);
                u_pos[dst][q] = u16(pos);
            }
        }
        __syncthreads();
        if (tid == 0) s_nu[src] = 0;
        __syncthreads();
    }
    // 3. 64 lanes per 8^3 node, one 2^3 block each: 4^3 child -> block -> the block's 8 voxel values, then its 4 half
    //    words of this plane.  A 4^3 or 2^3 leaf is spread over its blocks' lanes the same way.
    {
        const int fin = (ld - 3) & 1;
        const u32 n8 = s_nu[fin];
        for (u32 t = tid; t < n8 * 64; t += nthr) {
            const u32 e = t >> 6, c4 = (t >> 3) & 7, ci = t & 7;
            const u32 pos = u_pos[fin][e];
            const u64 id4 = __ldg(&children[size_t(u_idx[fin][e]) * 8 + c4]);
            if (id4 == 0) continue;
            T v[8];
            T uni = T(0);                                       // != 0: the block is one uniform 2^3 cube of that value
            bool voxels = false;                                // v[] holds the block's 8 voxel values
            if (id_is_leaf(id4)) {                              // the whole 4^3 node is one leaf
                uni = values[id_index(id4)];
            } else {
                const u64 bid = __ldg(&children[size_t(id_index(id4)) * 8 + ci]);
                if (bid == 0) continue;
                if (id_is_leaf(bid)) {
                    uni = values[id_index(bid)];
                } else {                                        // depth ld - 1: its children are the voxels
                    const uint4* row = reinterpret_cast<const uint4*>(&children[size_t(id_index(bid)) * 8]);
                    u64 ch[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 r = __ldg(row + i);
                        ch[2 * i] = u64(r.x) | (u64(r.y) << 32);
                        ch[2 * i + 1] = u64(r.z) | (u64(r.w) << 32);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = ch[i] != 0 ? values[id_index(ch[i])] : T(0);
                    voxels = true;
                }
            }
            const u32 x = ((pos & 15) << 2) + 4 * (c4 & 1) + 2 * (ci & 1),
                      y = (((pos >> 4) & 15) << 2) + 4 * ((c4 >> 1) & 1) + 2 * ((ci >> 1) & 1),
                      z = (((pos >> 8) & 15) << 2) + 4 * (c4 >> 2) + 2 * (ci >> 2);
            if (uni != T(0)) {                                  // same code path as a mixed block: the warp stays together
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = uni;
                voxels = true;
            }
            if (!voxels) continue;                              // a leaf holding the default value: nothing to set
            occ_block_words<T>(plane, x, y, z, v, [&](T val, u32 i, u32 bits) {
                const int slot = slot_of(val);
                if (slot < 0) return;
                if (plane == 0) count(slot, u32(__popc(bits)));
                atomicOr(&planes[size_t(slot) * OCC_HALVES + i], bits);
            });
        }
    }
    if (plane == 0) {                                            // voxel counts: one shared atomic per warp and slot
#pragma unroll
        for (int k = 0; k < OCC_MS_MAX; ++k) {
            u32 c = cnt[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
            if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[k], c);
        }
    }
    __syncthreads();
    if (s_over) {                                               // too many materials for shared memory
        if (tid == 0 && plane == 0) {
            overflow[b] = 1;
            atomicAdd(overflow_count, 1u);
        }
        return;
    }
    {
        const u32 qn = min(s_qn, u32(OCC_QUEUE));
        const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
        for (u32 q = warp; q < qn; q += nwarp) {
            const u32 e = s_queue[q];
            u32* pmat = planes + size_t(e >> 21) * OCC_HALVES;
            occ_region_words(plane, e & 63, (e >> 6) & 63, (e >> 12) & 63, (e >> 18) & 7, lane, 32,
                             [&](u32 i, u32 bits) { atomicOr(&pmat[i], bits); });
        }
    }
    if (tid == 0) {                                             // build(): materials in id order
        int n = 0;
        for (int k = 0; k < ms; ++k) n += s_mat[k] != 0;
        for (int k = 0; k < n; ++k) {
            const u64 id = occ_material_of<T>(T(s_mat[k]));
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += occ_material_of<T>(T(s_mat[j])) < id;
            s_order[rank] = k;
        }
        s_n = n;
        if (plane == 0) {
            overflow[b] = 0;
            n_materials[b] = u32(n);
        }
    }
    __syncthreads();
    const int n = s_n;
    if (plane == 0 && tid < n) {
        const int k = s_order[tid];
        material_ids[size_t(b) * max_materials + tid] = occ_material_of<T>(T(s_mat[k]));
        material_counts[size_t(b) * max_materials + tid] = s_cnt[k];
    }
    // stream the planes out, 16 bytes per thread and step: the materials in id order, and their OR = the global plane
    u32 alo = 0, ahi = 0;
    uint4* gout = reinterpret_cast<uint4*>(global + size_t(b) * OCC_ALL + size_t(plane) * OCC_PLANE);
    for (int i = tid; i < OCC_HALVES / 4; i += nthr) {
        uint4 g = make_uint4(0, 0, 0, 0);
        for (int r = 0; r < n; ++r) {
            const uint4 v = occ_smem[size_t(s_order[r]) * (OCC_HALVES / 4) + i];
            reinterpret_cast<uint4*>(per_material + (size_t(b) * max_materials + r) * OCC_ALL +
                                     size_t(plane) * OCC_PLANE)[i] = v;
            g.x |= v.x, g.y |= v.y, g.z |= v.z, g.w |= v.w;
        }
        gout[i] = g;
        alo |= g.x | g.z;
        ahi |= g.y | g.w;
    }
    // global_active (mesh.rs:451-461): the OR of every word of this plane = the mask of its bit axis
#pragma unroll
    for (int k = 16; k; k >>= 1) alo |= __shfl_xor_sync(FULL, alo, k), ahi |= __shfl_xor_sync(FULL, ahi, k);
    if ((tid & 31) == 0) {
        if (alo) atomicOr(&s_act[0], alo);
        if (ahi) atomicOr(&s_act[1], ahi);
    }
    __syncthreads();
    if (tid == 0) {
        const u64 t = u64(s_act[0]) | (u64(s_act[1]) << 32);
        const int i0 = plane == 0 ? 3 : plane == 1 ? 0 : 1;
        const int i1 = plane == 0 ? 5 : plane == 1 ? 4 : 2;
        active[size_t(b) * 6 + i0] = t;
        active[size_t(b) * 6 + i1] = t;
    }
}

}  // namespace vx

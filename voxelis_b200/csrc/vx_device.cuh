// vx_device.cuh — device-side vocabulary of the GPU-resident VoxInterner (sm_100a).
//
// Reference being replaced (paths under /root/reference/voxelis/src):
//   BlockId packing            core/block_id.rs:18-30,216-226
//   VoxInterner SoA pools      interner/mod.rs:25-40  (ref_counts, generations, children, values, hashes)
//   pattern maps               interner/hash.rs:15-38 (HashMap<u64,BlockId>, one for branches, one for leaves)
//
// Layout in HBM (one set per interner, all cudaMalloc'd once by vx_interner_create):
//   children[cap][8] u64   64 B per node, 64 B aligned  -> one 8-lane group reads/writes a node coalesced
//   values[cap]      T     leaf value / branch LOD value (mode of child values, core/voxel.rs:96-141)
//   refs[cap]        u32   in-degree from live unique branches + root handles
//   gens[cap]        u16   generation of the slot (bumped on recycle)
//   hashes[cap]      u64   64-bit key hash (table maintenance / future re-hash)
//   slots[nb][8]     u64   open-addressing BRANCH table, bucket = 8 slots = 64 B (2 sectors),
//                          slot = fp17<<47 | gen15<<32 | index32 ; 0 = empty
//   leaf_u8[256]     u64   u8 leaves: direct-mapped value -> BlockId (0 = absent)
//   leaf_keys/ids[]  u64   wider T: open-addressing leaf table, key = value | 1<<32
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace vx {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using ull = unsigned long long;

constexpr u32 FULL = 0xFFFFFFFFu;
constexpr u32 IDX_PENDING = 0xFFFFFFFFu;  // slot claimed, payload not yet published
constexpr u32 IDX_TOMB = 0xFFFFFFFEu;     // deleted slot (probe continues past it)
constexpr u64 ID_LEAF_BIT = 1ull << 63;
constexpr u64 ID_GENIDX = 0x00007FFFFFFFFFFFull;  // generation<<32 | index
constexpr u64 ID_PENDING = ~0ull;

// Sticky device error word (first error wins).
enum : u32 { ERR_NONE = 0, ERR_OOM = 1, ERR_TABLE_FULL = 2, ERR_INTERNAL = 3 };

__host__ __device__ inline u64 id_branch(u64 genidx, u32 types, u32 mask) {
    return (u64(types) << 55) | (u64(mask) << 47) | (genidx & ID_GENIDX);
}
__host__ __device__ inline u64 id_leaf(u64 genidx) { return ID_LEAF_BIT | (genidx & ID_GENIDX); }
__host__ __device__ inline u32 id_index(u64 id) { return u32(id); }
__host__ __device__ inline bool id_is_leaf(u64 id) { return (id >> 63) != 0; }
__host__ __device__ inline bool id_is_empty(u64 id) { return id == 0; }

// Device counters behind InternerStats (interner/stats.rs).  hits = calls - misses.
struct Counters {
    ull leaf_calls;       // get_or_create_leaf invocations the reference would have made
    ull branch_calls;     // get_or_create_branch invocations
    ull leaf_misses;      // new leaves
    ull branch_misses;    // new branches
    ull collapsed;        // bump_collapsed_branches (voxtree.rs:888-889,1063-1064)
    ull probe_steps;      // diagnostic: bucket reads in the global table
    ull cache_hits_local; // diagnostic: lookups answered by the per-warp shared-memory caches
    ull recycled;         // nodes released (total_deallocations)
};

struct InternerDev {
    u64* slots;
    u32 bucket_mask;  // nbuckets - 1
    u64* children;
    void* values;
    u32* refs;
    u16* gens;
    u64* hashes;
    u32 capacity;
    u32* next_index;
    u32* free_list;   // LIFO of recycled indices (interner/macros.rs:1-41)
    u32* free_count;
    u64* leaf_u8;
    u64* leaf_keys;
    u64* leaf_ids;
    u32 leaf_mask;
    u32* error;
    Counters* ctr;
};

// ---- memory-model helpers -------------------------------------------------------------------
// Interner words are written by one SM and read by others inside the same kernel, so every access
// to them is a *strong* (L2-coherent) access: L1 may hold stale neighbours of a freshly published
// node.  Publication = payload stores, fence.acq_rel.gpu, relaxed store of the slot; readers reach
// the payload through an address dependency on the slot value.
__device__ __forceinline__ u64 ld_strong(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_strong(const u32* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_strong_u8(const u8* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_strong_u16(const u16* p) {
    u32 v;
    asm volatile("{ .reg .u16 t; ld.relaxed.gpu.global.u16 t, [%1]; cvt.u32.u16 %0, t; }"
                 : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 16-byte strong load (two u64): L2-coherent like the scalar forms
__device__ __forceinline__ void ld_strong_v2(const u64* p, u64* a, u64* b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(*a), "=l"(*b) : "l"(p) : "memory");
}
// Weak (L1-cacheable) forms for the FIRST look at a bucket / stored row: hot keys are asked for by every warp of the
// grid, and strong loads send all of them to the one L2 slice that holds the line.  What a weak read returns may be
// stale, so only what cannot be wrong is concluded from it:
//   * bucket: a final entry never changes while a build kernel runs; a stale "empty" is validated by the CAS at L2; a
//     stale "pending", a full bucket or a fingerprint match whose row differs conclude nothing — the next round looks
//     again with strong loads;
//   * stored row: a row is exactly two 32-byte sectors of its own, L1 fills sector by sector, and a row is only ever
//     read through a published slot, i.e. after its one and only write of this life — so a cached row is the final
//     row.  A row match is a hit; a mismatch concludes nothing (strong reads next round).  Belt and braces: dead
//     indices are given all-zero rows where that is cheap (pool creation, release, the sparse reset), and an all-zero
//     row equals no key.
// The level kernels drop the SM's L1 in their prologue (bulk_prologue), so nothing cached is older than the launch.
__device__ __forceinline__ void ld_weak_v2(const u64* p, u64* a, u64* b) {
    asm volatile("ld.global.ca.v2.u64 {%0, %1}, [%2];" : "=l"(*a), "=l"(*b) : "l"(p) : "memory");
}
__device__ __forceinline__ u64 ld_weak(const u64* p) {
    u64 v;
    asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_strong(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_strong(u32* p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// Publisher side only: orders the payload stores before the slot store.  Unlike fence.acq_rel it does
// not invalidate the SM's L1 (no CCTL.IVALL): the publisher acquires nothing.
__device__ __forceinline__ void fence_release_gpu() { asm volatile("fence.release.gpu;" ::: "memory"); }
// CTA-shared memo words (CtaSmem::leaf) are read and written by every warp of the CTA with no barrier
// in between: all writers store the same value and a reader takes either 0 (-> global table) or that
// value.  Relaxed .cta accesses make those morally-strong operations under the PTX memory model, so the
// (benign, idempotent) race is not a data race and the 8-byte word cannot be observed torn.
__device__ __forceinline__ u64 lds_relaxed(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"(u32(__cvta_generic_to_shared(p))) : "memory");
    return v;
}
__device__ __forceinline__ void sts_relaxed(u64* p, u64 v) {
    asm volatile("st.relaxed.cta.shared.u64 [%0], %1;" ::"r"(u32(__cvta_generic_to_shared(p))), "l"(v) : "memory");
}

// Streaming (read-once) batch loads: keep them out of L1 so the
// interner's table and pools keep the cache.
__device__ __forceinline__ u64 ld_stream_u64(const void* p) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_stream_u16(const void* p) {
    u32 v;
    asm volatile("{ .reg .u16 t; ld.global.nc.L1::no_allocate.u16 t, [%1]; cvt.u32.u16 %0, t; }"
                 : "=r"(v) : "l"(p));
    return v;
}

// ---- hashing ---------------------------------------------------------------------------------
// Any hash is admissible: hash values never leave the interner and full keys are always compared
// (the reference's FxHash-only keying, interner/hash.rs:15-38, is a latent aliasing bug).
__host__ __device__ inline u64 mix64(u64 x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
// Branch hash = finish( sum_i child_i * HC_i ): a multilinear hash with one odd 64-bit multiplier per
// child position — one multiply per child instead of a full avalanche, the sum is order-sensitive
// through the multipliers, and the final mix spreads the entropy into the bucket (low) and the
// fingerprint (high) bits.  Both probing schemes (8-lane group, thread-per-key) use this formula.
#define VX_HC0 0x9E3779B97F4A7C15ull
#define VX_HC1 0xBF58476D1CE4E5B9ull
#define VX_HC2 0x94D049BB133111EBull
#define VX_HC3 0xD6E8FEB86659FD93ull
#define VX_HC4 0xC2B2AE3D27D4EB4Full
#define VX_HC5 0x165667B19E3779F9ull
#define VX_HC6 0x27D4EB2F165667C5ull
#define VX_HC7 0xFF51AFD7ED558CCDull
// per-lane form: one constant-bank load instead of a select chain
__device__ __constant__ u64 VX_HC_TABLE[8] = {VX_HC0, VX_HC1, VX_HC2, VX_HC3, VX_HC4, VX_HC5, VX_HC6, VX_HC7};
__device__ __forceinline__ u64 child_hash_lane(u64 child, int li) { return child * VX_HC_TABLE[li]; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__host__ __device__ inline u64 child_mult(int i) {
    return i == 0 ? VX_HC0 : i == 1 ? VX_HC1 : i == 2 ? VX_HC2 : i == 3 ? VX_HC3 : i == 4 ? VX_HC4
         : i == 5 ? VX_HC5 : i == 6 ? VX_HC6 : VX_HC7;
}
__host__ __device__ inline u64 child_hash(u64 child, int i) { return child * child_mult(i); }
__host__ __device__ inline u64 finish_hash(u64 h) {
    h ^= h >> 32;
    h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 29;
    h *= 0x94d049bb133111ebull;
    h ^= h >> 32;
    return h;
}
__host__ __device__ inline u64 leaf_hash(u64 v) { return mix64(v ^ 0x5851F42D4C957F2Dull); }

__device__ __forceinline__ void set_error(const InternerDev& in, u32 code) { atomicCAS(in.error, 0u, code); }

}  // namespace vx

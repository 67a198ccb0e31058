// vx_build.cuh — fresh-tree apply_batch(es) kernels (sm_100a).
//
// Replaces, for trees whose root is EMPTY (the north-star path):
//   VoxTree::apply_batch                 spatial/voxtree.rs:303-328
//   set_batch_at_depth_iterative         spatial/voxtree.rs:724-1118   (phases 0 / 1 / 2)
//   VoxInterner::get_or_create_leaf      interner/mod.rs:627-710
//   VoxInterner::get_or_create_branch    interner/mod.rs:716-829  (+ calc_average core/voxel.rs:96-141)
//
// What the reference's three phases reduce to on a fresh tree: the result is the
// canonical bottom-up DAG of the batch's effective volume,
//     block (depth D-1)  = EMPTY | Leaf(v) if set_mask==0xFF and 8 equal values | Branch(8 leaves/EMPTY)
//     parent             = absent if no child entered `paths`
//                        | that Leaf if types==0xFF and 8 identical ids (uniform collapse, :1050)
//                        | Branch(children) interned on the 8 child ids (level-agnostic table)
// so every level is a pure function of the level below and the work maps onto warps:
//
//   * a warp owns a Morton-contiguous run of <= 512 blocks (a 16^3-voxel sub-cube; Batch arrays are
//     Morton ordered, core/batch.rs:153-157).  Lane = block at depth D-1: one 8-byte (u8) or
//     2x16-byte (i32) coalesced load per lane, collapse / fill / skip decisions in registers.
//   * eight consecutive lanes are siblings, so the parent key (8 child ids) already sits one child
//     per lane: hashing is a 3-step shuffle reduction, key comparison is one ballot, and the node's
//     64-byte children row is read / written coalesced by the 8-lane group.
//   * duplicate detection runs warp -> shared memory -> HBM: __match_any_sync on the block value
//     word, per-warp direct-mapped caches in shared memory (block word -> id, children[8] -> id),
//     and only then the global open-addressing table (bucket of 8 slots = 64 B, probed by the
//     8-lane group with one ballot).
//   * refcount = in-degree: +1 per child slot only when a NEW unique branch is created, +1 per
//     tree root; u8 leaf increments are histogrammed in shared memory and flushed once per CTA.
//
// All warp collectives below execute in warp-uniform control flow; per-group work is predicated.
#pragma once
#include "vx_device.cuh"

#include <cstddef>

namespace vx {

constexpr int WARPS_PER_CTA = 8;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr int UNIT_BLOCKS = 512;  // blocks per warp work unit
// Per-warp cache sizes.  They set the CTA's shared memory (38 KiB at 128 / 16), and with three CTAs per SM that sets
// the SM's carve-out and so what is left as L1: 256 / 32 entries (62 KiB per CTA, 60 KiB of L1) built the repetitive
// worlds 6-14 % slower for the same hit rates; 64 / 16 and below lose hits on terrain (profiles/README.md, round 2).
#ifndef VX_UC
#define VX_UC 16
#endif
#ifndef VX_BC_U8
#define VX_BC_U8 128
#endif
constexpr int UC = VX_UC;         // per-warp parent-cache entries (children[8] -> id)

#define VX_FLAG_FILL 1u
#define VX_FLAG_PATCHES 2u

// ------------------------------------------------------------------------------------------------
// Value-type traits: how a lane holds the 8 values of its block.
// ------------------------------------------------------------------------------------------------
template <class T>
struct VT;

__device__ __forceinline__ u32 nzbytes(u64 x) {  // bit i set iff byte i != 0
    u64 y = (((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x) & 0x8080808080808080ull;
    return u32(((y >> 7) * 0x0102040810204080ull) >> 56);
}
__device__ __forceinline__ u64 expand_bits(u32 m) {  // bit i -> byte i = 0xFF
    u64 x = (u64(m) * 0x0101010101010101ull) & 0x8040201008040201ull;
    x = ((x + 0x7F7F7F7F7F7F7F7Full) & 0x8080808080808080ull) >> 7;
    return x * 0xFFull;
}

template <>
struct VT<u8> {
    static constexpr int KW = 1;
    static constexpr int BC = VX_BC_U8;  // per-warp block-cache entries
    struct Key {
        u64 w[1];
    };
    static __device__ __forceinline__ Key load(const void* values, size_t block) {
        Key k;
        k.w[0] = ld_stream_u64((const u8*)values + block * 8);
        return k;
    }
    static __device__ __forceinline__ Key zero() { return Key{{0}}; }
    static __device__ __forceinline__ u32 nz(const Key& k) { return nzbytes(k.w[0]); }
    static __device__ __forceinline__ u32 first(const Key& k) { return u32(k.w[0] & 0xFF); }
    static __device__ __forceinline__ u32 get(const Key& k, int i) { return u32(k.w[0] >> (8 * i)) & 0xFF; }
    static __device__ __forceinline__ bool uniform(const Key& k) {
        return k.w[0] == (k.w[0] & 0xFF) * 0x0101010101010101ull;
    }
    static __device__ __forceinline__ u32 ne_mask(const Key& k, u32 f) {  // bits where value != f
        return nzbytes(k.w[0] ^ (u64(f) * 0x0101010101010101ull));
    }
    static __device__ __forceinline__ Key select(const Key& k, u32 m, u32 f) {  // set ? value : f
        u64 bm = expand_bits(m);
        return Key{{(k.w[0] & bm) | ((u64(f) * 0x0101010101010101ull) & ~bm)}};
    }
    static __device__ __forceinline__ Key splat(u32 f) { return Key{{u64(f & 0xFF) * 0x0101010101010101ull}}; }
    static __device__ __forceinline__ u32 ne_key(const Key& a, const Key& b) { return nzbytes(a.w[0] ^ b.w[0]); }
    static __device__ __forceinline__ Key select_key(const Key& k, u32 m, const Key& o) {  // set ? value : old
        u64 bm = expand_bits(m);
        return Key{{(k.w[0] & bm) | (o.w[0] & ~bm)}};
    }
    // lane li of an 8-lane group contributes value v; every lane of the group receives the assembled key
    static __device__ __forceinline__ Key assemble(u32 v, int li) {
        u64 w = u64(v & 0xFF) << (8 * li);
        w |= __shfl_xor_sync(FULL, w, 1);
        w |= __shfl_xor_sync(FULL, w, 2);
        w |= __shfl_xor_sync(FULL, w, 4);
        return Key{{w}};
    }
    static __device__ __forceinline__ bool eq(const Key& a, const Key& b) { return a.w[0] == b.w[0]; }
    static __device__ __forceinline__ u32 hash(const Key& k) {  // index of the per-warp block cache
        return ((u32(k.w[0]) * 0x9E3779B1u) ^ (u32(k.w[0] >> 32) * 0x85EBCA77u)) >> 20;
    }
    // value of child `li` of the key held by lane `src` (full-warp shuffle)
    static __device__ __forceinline__ u32 bcast_value(const Key& k, int src, int li) {
        u64 w = __shfl_sync(FULL, k.w[0], src);
        return u32(w >> (8 * li)) & 0xFF;
    }
    static __device__ __forceinline__ Key bcast(const Key& k, int src) {
        return Key{{__shfl_sync(FULL, k.w[0], src)}};
    }
};

template <>
struct VT<int32_t> {
    static constexpr int KW = 4;
    static constexpr int BC = 32;
    struct Key {
        u64 w[4];
    };
    static __device__ __forceinline__ Key load(const void* values, size_t block) {
        const uint4* p = (const uint4*)((const u8*)values + block * 32);
        uint4 a = ld_stream_v4(p), b = ld_stream_v4(p + 1);
        Key k;
        k.w[0] = u64(a.x) | (u64(a.y) << 32);
        k.w[1] = u64(a.z) | (u64(a.w) << 32);
        k.w[2] = u64(b.x) | (u64(b.y) << 32);
        k.w[3] = u64(b.z) | (u64(b.w) << 32);
        return k;
    }
    static __device__ __forceinline__ Key zero() { return Key{{0, 0, 0, 0}}; }
    static __device__ __forceinline__ u32 get(const Key& k, int i) { return u32(k.w[i >> 1] >> (32 * (i & 1))); }
    static __device__ __forceinline__ u32 nz(const Key& k) {
        u32 m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= u32(get(k, i) != 0) << i;
        return m;
    }
    static __device__ __forceinline__ u32 first(const Key& k) { return u32(k.w[0]); }
    static __device__ __forceinline__ bool uniform(const Key& k) {
        u64 f = u64(u32(k.w[0])) * 0x0000000100000001ull;
        return k.w[0] == f && k.w[1] == f && k.w[2] == f && k.w[3] == f;
    }
    static __device__ __forceinline__ u32 ne_mask(const Key& k, u32 f) {
        u32 m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= u32(get(k, i) != f) << i;
        return m;
    }
    static __device__ __forceinline__ Key select(const Key& k, u32 m, u32 f) {
        Key r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u32 lo = (m >> (2 * j)) & 1 ? u32(k.w[j]) : f;
            u32 hi = (m >> (2 * j + 1)) & 1 ? u32(k.w[j] >> 32) : f;
            r.w[j] = u64(lo) | (u64(hi) << 32);
        }
        return r;
    }
    static __device__ __forceinline__ Key splat(u32 f) {
        u64 w = u64(f) * 0x0000000100000001ull;
        return Key{{w, w, w, w}};
    }
    static __device__ __forceinline__ u32 ne_key(const Key& a, const Key& b) {
        u32 m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= u32(get(a, i) != get(b, i)) << i;
        return m;
    }
    static __device__ __forceinline__ Key select_key(const Key& k, u32 m, const Key& o) {
        Key r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u32 lo = (m >> (2 * j)) & 1 ? u32(k.w[j]) : u32(o.w[j]);
            u32 hi = (m >> (2 * j + 1)) & 1 ? u32(k.w[j] >> 32) : u32(o.w[j] >> 32);
            r.w[j] = u64(lo) | (u64(hi) << 32);
        }
        return r;
    }
    static __device__ __forceinline__ Key assemble(u32 v, int li) {
        Key r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u64 w = (li >> 1) == j ? u64(v) << (32 * (li & 1)) : 0;
            w |= __shfl_xor_sync(FULL, w, 1);
            w |= __shfl_xor_sync(FULL, w, 2);
            w |= __shfl_xor_sync(FULL, w, 4);
            r.w[j] = w;
        }
        return r;
    }
    static __device__ __forceinline__ bool eq(const Key& a, const Key& b) {
        return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3];
    }
    static __device__ __forceinline__ u32 hash(const Key& k) {
        return u32(mix64(k.w[0] + mix64(k.w[1] + mix64(k.w[2] + mix64(k.w[3])))) >> 32);
    }
    static __device__ __forceinline__ Key bcast(const Key& k, int src) {
        Key r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.w[j] = __shfl_sync(FULL, k.w[j], src);
        return r;
    }
    static __device__ __forceinline__ u32 bcast_value(const Key& k, int src, int li) {
        Key r = bcast(k, src);
        u64 w = (li >> 1) == 0 ? r.w[0] : (li >> 1) == 1 ? r.w[1] : (li >> 1) == 2 ? r.w[2] : r.w[3];
        return u32(w >> (32 * (li & 1)));
    }
};

// ------------------------------------------------------------------------------------------------
// Shared memory
// ------------------------------------------------------------------------------------------------
constexpr int LC = 64;  // per-warp leaf cache entries (wide T only)

template <class T>
struct WarpSmem {
    u64 bkey[VT<T>::BC * VT<T>::KW];  // block value words -> id
    u64 bval[VT<T>::BC];
    u64 ukey[UC * 8];  // children[8] -> id
    u64 uval[UC];
    u64 l1[64];       // nodes being joined (level-1 parents of a unit; unit / cube nodes during joins)
    u8 p1[64];        // "entered paths" flag per node
    u64 run_ids[8];   // run == 8: the unit nodes of the cube this warp owns
    u8 run_pres[8];
    u32 lkey[sizeof(T) == 1 ? 1 : LC];  // wide T only: value -> leaf id
    u64 lid[sizeof(T) == 1 ? 1 : LC];
};
constexpr int RH = 512;  // entries of the CTA's in-degree combining table
struct CtaSmem {
    u64 leaf[256];     // u8: value -> leaf id (lazy copy of InternerDev::leaf_u8)
    u32 leafref[256];  // u8: pending in-degree increments of leaves
    u32 rkey[RH];      // pending in-degree increments of other children: node index + 1 (0 = free) ...
    u32 rcnt[RH];      // ... and how many; hot children (flat ground, solid rock) would otherwise take
                       // thousands of same-address RED.ADDs that serialise in one L2 slice
    u32 warps_done;    // exit arrival counter (the last warp flushes leafref / rkey,rcnt)
    u32 memo_stat[2];  // bulk builder's unit memo: units of this CTA that were aliased / that were new (vx_bulk.cuh)
};

struct Tally {  // per-lane statistics, reduced once at kernel exit
    u32 leaf_calls = 0, branch_calls = 0, leaf_miss = 0, branch_miss = 0, collapsed = 0, probes = 0, local = 0;
};

template <class T>
struct Ctx {
    InternerDev in;
    WarpSmem<T>* ws;
    CtaSmem* cs;
    int lane, li, gs;  // lane, lane within 8-group, first lane of my group
    bool use_free;     // the free list holds recycled indices: pop them before next_index (macros.rs:1-41)
    bool tpk_only;     // block level: always thread-per-key (bulk builder); the fused kernel uses the hybrid
    bool weak_first;   // first look at a bucket / row through L1 (hot keys are read by every warp of the grid; vx_device.cuh)
    Tally t;
};

// inc_ref of a child of a NEW node (interner/mod.rs:301-330 via inc_all_child_refs): combined per CTA in
// shared memory, flushed once at kernel exit; falls back to the global counter when the two probed
// entries belong to other nodes.  Nothing reads refs while an apply kernel runs.
template <class T>
__device__ __forceinline__ void ref_add(Ctx<T>& c, u32 idx) {
    u32 h = (idx * 0x9E3779B1u) >> 23;  // RH = 512
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        u32 k = ((volatile u32*)c.cs->rkey)[h];
        if (k == 0) {
            k = atomicCAS(&c.cs->rkey[h], 0u, idx + 1);
            if (k == 0) k = idx + 1;
        }
        if (k == idx + 1) {
            atomicAdd(&c.cs->rcnt[h], 1u);
            return;
        }
        h = (h + 1) & (RH - 1);
    }
    atomicAdd(&c.in.refs[idx], 1u);
}

// one index allocation for n new nodes (aggregated per warp step).  Measured and dropped: 64 striped cursors instead of
// this single counter — 2.68 ms against 2.69 on the all-miss input: the wait on this atomic is its round trip, not
// same-address throughput (profiles/README.md).
__device__ __forceinline__ u32 alloc_n(const InternerDev& in, u32 n) { return atomicAdd(in.next_index, n); }

// get_next_index_macro! (interner/macros.rs:1-41) for one node: recycled index first (LIFO), else
// next_index++.  `*gen` = generation to stamp into the BlockId.  Indices >= capacity mean "Out of memory".
__device__ __forceinline__ u32 alloc_one(const InternerDev& in, bool use_free, u32* gen) {
    *gen = 0;
    if (use_free) {
        int old = atomicSub((int*)in.free_count, 1);
        if (old > 0) {
            u32 idx = ld_strong(&in.free_list[old - 1]);
            *gen = ld_strong_u16(&in.gens[idx]);
            return idx;
        }
    }
    return atomicAdd(in.next_index, 1u);
}

// ------------------------------------------------------------------------------------------------
// Leaves — get_or_create_leaf (interner/mod.rs:627-710).  Warp-converged; `need` predicates lanes.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void leaf_payload(const InternerDev& in, u32 idx, u32 v) {
    ((T*)in.values)[idx] = T(v);
    in.hashes[idx] = leaf_hash(v);
    // "no children" is how every later pass tells a leaf from a branch (release, rehash, heights, VTM, reset): the row
    // may still hold the ids of a branch that lived at this index before the last vx_interner_reset
    ulonglong2* row = reinterpret_cast<ulonglong2*>(in.children + size_t(idx) * 8);
#pragma unroll
    for (int k = 0; k < 4; ++k) row[k] = make_ulonglong2(0, 0);
}

__device__ inline u64 leaf_get(Ctx<u8>& c, u32 v, bool need) {
    need = need && v != 0;
    u64 id = need ? lds_relaxed(&c.cs->leaf[v]) : 0;
    bool miss = need && id == 0;
    bool first = c.weak_first;  // every CTA of the grid asks for the same few leaves at its start: first look through L1
    while (__any_sync(FULL, miss)) {
        if (miss) {
            u64 g = first ? ld_weak(&c.in.leaf_u8[v]) : ld_strong(&c.in.leaf_u8[v]);
            first = false;
            if (g == 0) {
                u64 old = atomicCAS((ull*)&c.in.leaf_u8[v], 0ull, (ull)ID_PENDING);
                if (old == 0) {  // this lane creates the leaf
                    u32 gen;
                    u32 idx = alloc_one(c.in, c.use_free, &gen);
                    if (idx >= c.in.capacity) {
                        set_error(c.in, ERR_OOM);
                        st_strong(&c.in.leaf_u8[v], 0);
                        miss = false;
                    } else {
                        leaf_payload<u8>(c.in, idx, v);
                        id = id_leaf((u64(gen) << 32) | idx);
                        fence_release_gpu();
                        st_strong(&c.in.leaf_u8[v], id);
                        sts_relaxed(&c.cs->leaf[v], id);
                        c.t.leaf_miss++;
                        miss = false;
                    }
                }
            } else if (g != ID_PENDING) {
                id = g;
                sts_relaxed(&c.cs->leaf[v], g);
                miss = false;
            }
        }
    }
    return id;
}

__device__ inline u64 leaf_get(Ctx<int32_t>& c, u32 v, bool need) {
    need = need && v != 0;
    u32 e = (v * 0x9E3779B1u) >> 26;  // LC = 64
    u64 id = 0;
    bool miss = need;
    if (need && c.ws->lkey[e] == v) {
        id = c.ws->lid[e];
        miss = id == 0;
    }
    bool from_global = miss;
    u64 mykey = u64(v) | (1ull << 32);
    u32 s = u32(leaf_hash(v)) & c.in.leaf_mask;
    int guard = 0;
    u32 spins = 0;
    while (__any_sync(FULL, miss)) {
        if (miss) {
            u64 k = ld_strong(&c.in.leaf_keys[s]);
            if (k == 0) {
                u64 old = atomicCAS((ull*)&c.in.leaf_keys[s], 0ull, (ull)mykey);
                if (old == 0) {
                    u32 gen;
                    u32 idx = alloc_one(c.in, c.use_free, &gen);
                    if (idx >= c.in.capacity) {
                        // "Out of memory" (interner/macros.rs:38).  The key stays claimed (probes may have walked past
                        // it), so the id word tells everyone waiting for this value that it will never come
                        set_error(c.in, ERR_OOM);
                        st_strong(&c.in.leaf_ids[s], ID_PENDING);
                        miss = false;
                    } else {
                        leaf_payload<int32_t>(c.in, idx, v);
                        id = id_leaf((u64(gen) << 32) | idx);
                        fence_release_gpu();
                        st_strong(&c.in.leaf_ids[s], id);
                        c.t.leaf_miss++;
                        miss = false;
                    }
                } else if (old != mykey) {
                    s = (s + 1) & c.in.leaf_mask;
                }
            } else if (k == mykey) {
                u64 g = ld_strong(&c.in.leaf_ids[s]);
                if (g != 0) {
                    id = g == ID_PENDING ? 0 : g;  // ID_PENDING: its creator ran out of memory (the interner is poisoned)
                    miss = false;
                } else if (((++spins) & 0xFFF) == 0 && ld_strong(c.in.error) != ERR_NONE) {
                    miss = false;  // whoever claimed the key may have left the kernel on an error: do not wait for ever
                }
            } else {
                s = (s + 1) & c.in.leaf_mask;
                if (++guard > (1 << 22)) {
                    set_error(c.in, ERR_TABLE_FULL);
                    miss = false;
                }
            }
        }
    }
    // refresh the per-warp cache; one writer per entry (match on e) keeps key/id pairs untorn
    u32 wmask = __ballot_sync(FULL, from_global && id != 0);
    __syncwarp();
    if (from_global && id != 0) {
        u32 same = __match_any_sync(wmask, e);
        if ((__ffs(same) - 1) == c.lane) {
            c.ws->lkey[e] = v;
            c.ws->lid[e] = id;
        }
    }
    __syncwarp();
    return id;
}

template <class T>
__device__ __forceinline__ u32 child_value(const InternerDev& in, u64 child);
template <>
__device__ __forceinline__ u32 child_value<u8>(const InternerDev& in, u64 child) {
    return ld_strong_u8((const u8*)in.values + id_index(child));
}
template <>
__device__ __forceinline__ u32 child_value<int32_t>(const InternerDev& in, u64 child) {
    return ld_strong((const u32*)in.values + id_index(child));
}

// ------------------------------------------------------------------------------------------------
// Branches — get_or_create_branch (interner/mod.rs:716-829), upper levels: up to four keys per warp,
// one per 8-lane group, child i of the key in lane gs+i.  Returns the branch id to every lane of a
// group with need == true.
// ------------------------------------------------------------------------------------------------
// `block_level`: the children are voxels (Leaf(cval) / EMPTY) — used for iterations where at most four
// distinct block keys miss the per-warp cache; busier iterations go through intern_block below.
template <class T>
#ifdef VX_INTERN_NOINLINE
__device__ __noinline__
#else
__device__ inline
#endif
u64 intern_branch(Ctx<T>& c, bool need, u64 child, bool block_level, u32 cval) {
    const int li = c.li, gs = c.gs, lane = c.lane;
    const InternerDev& in = c.in;
    if (!__any_sync(FULL, need)) return 0;
    // types / mask are functions of the children (voxtree.rs:846-863, :959-1000)
    u32 leafb = (__ballot_sync(FULL, id_is_leaf(child)) >> gs) & 0xFF;
    u32 presb = (__ballot_sync(FULL, child != 0) >> gs) & 0xFF;
    u64 h = child_hash_lane(child, li);
    h += __shfl_xor_sync(FULL, h, 1);
    h += __shfl_xor_sync(FULL, h, 2);
    h += __shfl_xor_sync(FULL, h, 4);
    h = finish_hash(h);
    const u32 fp = u32(h >> 47);
    u32 bucket = u32(h) & in.bucket_mask;

    u64 result = 0;
    bool done = !need;
    // ---- per-warp parent cache (the block level has its own value-keyed cache)
    const u32 ue = u32(h >> 32) & (UC - 1);
    if (!block_level) {
        u64 ck = c.ws->ukey[ue * 8 + li];
        u32 eqb = (__ballot_sync(FULL, need && ck == child) >> gs) & 0xFF;
        if (need && eqb == 0xFF) {
            result = c.ws->uval[ue];
            done = true;
            if (li == 0) c.t.local++;
        }
    }
    const bool went_global = !done;
    u32 skip = 0;
    int guard = 0;
    bool first = c.weak_first;
    while (__any_sync(FULL, !done)) {
        u64 slot = 0;
        if (!done) slot = first ? ld_weak(&in.slots[size_t(bucket) * 8 + li]) : ld_strong(&in.slots[size_t(bucket) * 8 + li]);
        const u32 lo = u32(slot);
        bool fpm = !done && slot != 0 && lo != IDX_TOMB && u32(slot >> 47) == fp && !((skip >> li) & 1);
        u32 mb = (__ballot_sync(FULL, fpm) >> gs) & 0xFF;
        u32 eb = (__ballot_sync(FULL, !done && slot == 0) >> gs) & 0xFF;
        if (!done && li == 0) c.t.probes++;
        // -- candidate with matching fingerprint: compare the stored children row
        const bool has_cand = mb != 0;
        const int k = has_cand ? (__ffs(mb) - 1) : 0;
        u64 cslot = __shfl_sync(FULL, slot, gs + k);
        const bool cand_pending = u32(cslot) == IDX_PENDING;
        const bool do_cmp = !done && has_cand && !cand_pending;
        u64 stored = 0;
        if (do_cmp) stored = first ? ld_weak(&in.children[size_t(u32(cslot)) * 8 + li]) : ld_strong(&in.children[size_t(u32(cslot)) * 8 + li]);
        u32 eqb = (__ballot_sync(FULL, do_cmp && stored == child) >> gs) & 0xFF;
        if (do_cmp) {
            if (eqb == 0xFF) {
                result = id_branch(cslot, leafb, presb);  // hit
                done = true;
            } else if (!first) {
                skip |= 1u << k;
            }  // a mismatch seen through L1 proves nothing (the row may be stale): the next round looks again at L2
        }
        // -- no candidate: claim the first empty slot of the bucket, or move on
        const bool want_claim = !done && !has_cand && eb != 0;
        const int ek = want_claim ? (__ffs(eb) - 1) : 0;
        bool claimed = false;
        if (want_claim && li == 0) {
            u64 old = atomicCAS((ull*)&in.slots[size_t(bucket) * 8 + ek], 0ull, (ull)((u64(fp) << 47) | IDX_PENDING));
            claimed = old == 0;
        }
        const u32 cb = __ballot_sync(FULL, claimed);  // bits at lanes 0, 8, 16, 24
        if (cb != 0) {                                // warp-uniform: someone creates a node
            const bool mine = (cb >> gs) & 1;
            u32 idx, gen = 0;
            if (!c.use_free) {  // one aggregated atomic for all nodes the warp creates this round
                u32 base = 0;
                if (lane == 0) base = alloc_n(in, u32(__popc(cb)));
                base = __shfl_sync(FULL, base, 0);
                idx = base + __popc(cb & ((1u << gs) - 1));
            } else {
                idx = 0;
                if (mine && li == 0) idx = alloc_one(in, true, &gen);
                idx = __shfl_sync(FULL, idx, gs);
                gen = __shfl_sync(FULL, gen, gs);
            }
            const u64 genidx = (u64(gen) << 32) | idx;
            const bool oom = idx >= in.capacity;
            // LOD value = mode of the child values (core/voxel.rs:96-141): most frequent; ties go to
            // a non-default value, then to the earliest first occurrence.
            u32 v = cval;
            if (!block_level) v = (mine && !oom && child != 0) ? child_value<T>(in, child) : 0;
            u32 mm = __match_any_sync(FULL, (u64(v) << 2) | u64(gs >> 3));
            u32 gm = (mm >> gs) & 0xFF;
            u32 score = (u32(__popc(gm)) << 8) | (u32(v != 0) << 7) | (u32(8 - __ffs(gm)) << 3) | u32(li);
            u32 best = score;
            best = max(best, __shfl_xor_sync(FULL, best, 1));
            best = max(best, __shfl_xor_sync(FULL, best, 2));
            best = max(best, __shfl_xor_sync(FULL, best, 4));
            u32 mode = __shfl_sync(FULL, v, gs + int(best & 7));  // the winner is a first occurrence
            if (mine) {
                if (oom) {
                    set_error(in, ERR_OOM);
                } else {
                    in.children[size_t(idx) * 8 + li] = child;
                    if (child != 0) {
                        if (block_level && sizeof(T) == 1)
                            atomicAdd(&c.cs->leafref[cval], 1u);
                        else
                            atomicAdd(&in.refs[id_index(child)], 1u);
                    }
                    if (li == 0) {
                        ((T*)in.values)[idx] = T(mode);
                        in.hashes[idx] = h;
                        c.t.branch_miss++;
                    }
                }
            }
            __syncwarp();
            if (mine) {
                if (li == 0) {
                    fence_release_gpu();
                    // out of memory: hand the slot back (the interner is poisoned, results are discarded)
                    st_strong(&in.slots[size_t(bucket) * 8 + ek], oom ? u64(0) : ((u64(fp) << 47) | genidx));
                }
                result = oom ? 0 : id_branch(genidx, leafb, presb);
                done = true;
            }
        }
        if (!done && !has_cand && eb == 0 && !first) {  // bucket full and no match: next bucket
            bucket = (bucket + 1) & in.bucket_mask;
            skip = 0;
            if (++guard > (1 << 22)) {
                set_error(in, ERR_TABLE_FULL);
                done = true;
            }
        }
        first = false;
    }
    if (!block_level) {
        // refresh the parent cache; among groups mapping to the same entry only the lowest writes,
        // so an entry is never a mix of two keys
        bool wr = went_global && result != 0;
        u32 wb = __ballot_sync(FULL, wr && li == 0);
        bool write = wr;
#pragma unroll
        for (int g2 = 0; g2 < 3; ++g2) {
            u32 e2 = __shfl_sync(FULL, ue, g2 * 8);
            if (((wb >> (g2 * 8)) & 1) && g2 * 8 < gs && e2 == ue) write = false;
        }
        __syncwarp();  // the entry's earlier reads by other lanes are done (WAR)
        if (write) {
            c.ws->ukey[ue * 8 + li] = child;
            if (li == 0) c.ws->uval[ue] = result;
        }
        __syncwarp();
    }
    return result;
}

// mode of eight values with the reference's tie-breaks (core/voxel.rs:96-141), one thread.  Only runs
// when a new block-level branch is created; kept out of line to protect the instruction cache.
__device__ __noinline__ u32 mode8(u32 v0, u32 v1, u32 v2, u32 v3, u32 v4, u32 v5, u32 v6, u32 v7) {
    const u32 v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
    u32 best = v[0], best_score = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u32 cnt = 0;
        bool first = true;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cnt += v[j] == v[i];
            if (j < i && v[j] == v[i]) first = false;
        }
        u32 score = (cnt << 1) | u32(v[i] != 0);  // count, then non-default; earlier first occurrence wins ties
        if (first && score > best_score) {
            best_score = score;
            best = v[i];
        }
    }
    return best;
}

// calc_average with at most ONE distinct non-default value among the eight (most terrain nodes): that value wins when
// it holds at least as many children as the default does — a tie goes to the non-default value (core/voxel.rs:128-136).
// Returns false when two different non-default values are present (-> mode8).
__device__ __forceinline__ bool mode8_two_valued(u32 v0, u32 v1, u32 v2, u32 v3, u32 v4, u32 v5, u32 v6, u32 v7, u32* out) {
    const u32 x = v0 ? v0 : v1 ? v1 : v2 ? v2 : v3 ? v3 : v4 ? v4 : v5 ? v5 : v6 ? v6 : v7;  // first non-default value
    const u32 v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
    u32 cnt = 0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        cnt += v[i] == x;
        ok = ok && (v[i] == x || v[i] == 0);
    }
    *out = (x != 0 && cnt >= 4) ? x : 0;   // x == 0: all default
    return ok;
}
__device__ __forceinline__ u32 lod_value(u32 v0, u32 v1, u32 v2, u32 v3, u32 v4, u32 v5, u32 v6, u32 v7) {
    u32 r;
    if (mode8_two_valued(v0, v1, v2, v3, v4, v5, v6, v7, &r)) return r;
    return mode8(v0, v1, v2, v3, v4, v5, v6, v7);
}

// Ids of the eight voxel children (Leaf(value) / EMPTY) of a block key.  u8: re-derived on demand from
// the CTA's value -> leaf table in shared memory (no registers held); wider T: held in registers.
template <class T>
struct ChildIds;
template <>
struct ChildIds<u8> {
    const u64* leaf;
    u64 w;
    __device__ __forceinline__ void init(Ctx<u8>& c, const VT<u8>::Key& eff, bool need) {
        leaf = c.cs->leaf;
        w = eff.w[0];
        bool missing = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            u32 v = u32(w >> (8 * i)) & 0xFF;
            missing = missing || (need && v != 0 && lds_relaxed(&leaf[v]) == 0);
        }
        if (__any_sync(FULL, missing)) {  // rare: some leaf does not exist yet (get_or_create_leaf)
#pragma unroll 1
            for (int i = 0; i < 8; ++i) leaf_get(c, u32(w >> (8 * i)) & 0xFF, need);
        }
    }
    __device__ __forceinline__ u64 get(int i) const {
        u32 v = u32(w >> (8 * i)) & 0xFF;
        return v ? lds_relaxed(&leaf[v]) : 0;
    }
};
template <>
struct ChildIds<int32_t> {
    u64 id[8];
    __device__ __forceinline__ void init(Ctx<int32_t>& c, const VT<int32_t>::Key& eff, bool need) {
#pragma unroll
        for (int i = 0; i < 8; ++i) id[i] = leaf_get(c, VT<int32_t>::get(eff, i), need);
    }
    __device__ __forceinline__ u64 get(int i) const { return id[i]; }
};

// ------------------------------------------------------------------------------------------------
// Branches at the BLOCK level (children are voxels): thread-per-key.  Every lane with need == true
// owns one key (its block's eight effective values) and probes for it on its own: the 64-byte bucket
// and the stored children row are read as 4 x 16-byte loads, so up to 32 lookups of a warp are in
// flight at once (the 8-lane scheme above has 4).  Inserts of one warp step share a single
// aggregated index allocation and a single fence.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ inline u64 intern_block(Ctx<T>& c, bool need, const typename VT<T>::Key& eff) {
    using V = VT<T>;
    const InternerDev& in = c.in;
    if (!__any_sync(FULL, need)) return 0;
    ChildIds<T> ch;
    ch.init(c, eff, need);
    u64 h = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) h += child_hash(ch.get(i), i);
    h = finish_hash(h);
    const u32 nzm = V::nz(eff);  // types == mask == voxels present (voxtree.rs:846-863)
    const u32 fp = u32(h >> 47);
    u32 bucket = u32(h) & in.bucket_mask;
    u64 result = 0;
    bool done = !need;
    u32 skip = 0;
    int guard = 0;
    bool first = c.weak_first;
    while (__any_sync(FULL, !done)) {
        bool claimed = false;
        int ek = 0;
        if (!done) {
            const u64* bp = &in.slots[size_t(bucket) * 8];
            c.t.probes++;
            // the whole 64-byte bucket first (four independent 16-byte loads in flight), then the scan
            u64 sl[8];
            if (first) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ld_weak_v2(bp + 2 * j, &sl[2 * j], &sl[2 * j + 1]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) ld_strong_v2(bp + 2 * j, &sl[2 * j], &sl[2 * j + 1]);
            }
            u32 mb = 0, eb = 0, pb = 0;
            u64 cand = 0;
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                const u64 sv = sl[k];
                const u32 lo = u32(sv);
                if (sv == 0)
                    eb |= 1u << k;
                else if (lo != IDX_TOMB && u32(sv >> 47) == fp && !((skip >> k) & 1)) {
                    if (lo == IDX_PENDING)
                        pb |= 1u << k;
                    else {
                        mb |= 1u << k;
                        cand = sv;  // ends up as the lowest matching slot
                    }
                }
            }
            if (mb) {  // compare the stored children row with the key (row loaded in one go as well)
                const u64* rp = &in.children[size_t(u32(cand)) * 8];
                u64 r[8];
                if (first) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ld_weak_v2(rp + 2 * j, &r[2 * j], &r[2 * j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ld_strong_v2(rp + 2 * j, &r[2 * j], &r[2 * j + 1]);
                }
                bool eq = true;
#pragma unroll
                for (int i = 0; i < 8; ++i) eq = eq && r[i] == ch.get(i);
                if (eq) {
                    result = id_branch(cand, nzm, nzm);
                    done = true;
                } else if (!first) {
                    skip |= 1u << (__ffs(mb) - 1);
                }  // a mismatch seen through L1 proves nothing: the next round looks again at L2
            } else if (pb) {
                // a slot with my fingerprint is being published (possibly my key): look again
            } else if (eb) {
                ek = __ffs(eb) - 1;
                u64 old = atomicCAS((ull*)&in.slots[size_t(bucket) * 8 + ek], 0ull, (ull)((u64(fp) << 47) | IDX_PENDING));
                claimed = old == 0;
            } else if (!first) {
                bucket = (bucket + 1) & in.bucket_mask;
                skip = 0;
                if (++guard > (1 << 22)) {
                    set_error(in, ERR_TABLE_FULL);
                    done = true;
                }
            }
        }
        // ---- nodes created in this step: one index allocation, one fence, then publish (inside the
        //      loop, so a lane waiting on a PENDING slot of its own warp always sees it resolve)
        const u32 cb = __ballot_sync(FULL, claimed);
        if (cb != 0) {
            u32 idx = 0, gen = 0;
            if (!c.use_free) {
                u32 base = 0;
                if (c.lane == 0) base = alloc_n(in, u32(__popc(cb)));
                base = __shfl_sync(FULL, base, 0);
                idx = base + __popc(cb & ((1u << c.lane) - 1));
            } else if (claimed) {
                idx = alloc_one(in, true, &gen);
            }
            const u64 genidx = (u64(gen) << 32) | idx;
            const bool oom = idx >= in.capacity;
            if (claimed) {
                if (oom) {
                    set_error(in, ERR_OOM);
                } else {
                    u64* rp = &in.children[size_t(idx) * 8];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<ulonglong2*>(rp + 2 * j) = make_ulonglong2(ch.get(2 * j), ch.get(2 * j + 1));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const u32 v = V::get(eff, i);
                        if (v != 0) {
                            if (sizeof(T) == 1)
                                atomicAdd(&c.cs->leafref[v], 1u);
                            else
                                atomicAdd(&in.refs[id_index(ch.get(i))], 1u);
                        }
                    }
                    ((T*)in.values)[idx] = T(lod_value(V::get(eff, 0), V::get(eff, 1), V::get(eff, 2), V::get(eff, 3),
                                                   V::get(eff, 4), V::get(eff, 5), V::get(eff, 6), V::get(eff, 7)));
                    in.hashes[idx] = h;
                    c.t.branch_miss++;
                }
                fence_release_gpu();
                // out of memory: hand the slot back (the interner is poisoned, results are discarded)
                st_strong(&in.slots[size_t(bucket) * 8 + ek], oom ? u64(0) : ((u64(fp) << 47) | genidx));
                result = oom ? 0 : id_branch(genidx, nzm, nzm);
                done = true;
            }
        }
        first = false;
    }
    return result;
}

// ------------------------------------------------------------------------------------------------
// Old-tree lookup for applies on NON-EMPTY trees (voxtree.rs:785-822, :930-952): the node the old
// tree holds at position `prefix` (d three-bit child indices, MSB first) of depth d.  Returns a
// branch id, the enclosing Leaf if the old tree is uniform above/at that position, or EMPTY.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 old_at(const InternerDev& in, u64 root, u32 prefix, int d) {
    u64 node = root;
    for (int k = 0; k < d; ++k) {
        if (node == 0 || id_is_leaf(node)) return node;
        u32 ci = (prefix >> (3 * (d - 1 - k))) & 7;
        node = ld_strong(&in.children[size_t(id_index(node)) * 8 + ci]);
    }
    return node;
}

// ------------------------------------------------------------------------------------------------
// One level of phase 2 (voxtree.rs:905-1106) for four parents per warp: lanes gs..gs+7 hold the
// eight child ids; children that did not enter `paths` already carry what the reference clones
// into the slot (old child / enclosing old leaf / fill leaf / EMPTY, :1013-1048).  `old_self` = the
// same for the parent itself, used when none of its children entered `paths`.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ inline u64 parent_node(Ctx<T>& c, u64 child, bool present, u64 old_self, bool* ppresent) {
    const int gs = c.gs;
    u64 c0 = __shfl_sync(FULL, child, gs);
    u32 same = (__ballot_sync(FULL, child == c0) >> gs) & 0xFF;
    u32 pres = (__ballot_sync(FULL, present) >> gs) & 0xFF;
    const bool any_present = pres != 0;
    // types == 0xFF and eight identical ids -> uniform collapse (:1050, :1062-1075)
    const bool collapse = same == 0xFF && id_is_leaf(c0);
    if (any_present && c.li == 0) {
        if (collapse)
            c.t.collapsed++;
        else
            c.t.branch_calls++;
    }
    u64 id = intern_branch<T>(c, any_present && !collapse, child, false, 0);
    if (any_present && collapse) id = c0;
    if (!any_present) id = old_self;
    *ppresent = any_present;
    return id;
}

// ------------------------------------------------------------------------------------------------
// Phase 1 (voxtree.rs:770-897) for 32 blocks per warp, one per lane.  `oldw` = the eight voxel values
// the block has before the batch (old tree / fill / 0), `old_id` = the node standing there.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ inline u64 block_node(Ctx<T>& c, bool active, typename VT<T>::Key vals, u32 set_mask,
                                 typename VT<T>::Key oldw, u64 old_id, bool* present) {
    using V = VT<T>;
    const int lane = c.lane;
    // a set bit whose value is the default cannot come from Batch::set (batch.rs:162-168): ignored
    u32 m = active ? (set_mask & V::nz(vals)) : 0;
    const bool all_same = m == 0xFF && V::uniform(vals);  // :826
    const u32 changed_bits = m & V::ne_key(vals, oldw);   // :853-856 unchanged voxels are skipped
    const bool touched = all_same || changed_bits != 0;   // :865-868
    *present = touched;
    const bool need_branch = touched && !all_same;
    typename V::Key eff = V::select_key(vals, m, oldw);
    if (touched) {
        c.t.leaf_calls += all_same ? 1u : u32(__popc(changed_bits));
        if (all_same)
            c.t.collapsed++;  // :888-889
        else
            c.t.branch_calls++;
    }
    u64 id = touched ? 0 : old_id;  // untouched block: whatever stands there already
    // uniform collapse -> Leaf(value)
    u64 lid = leaf_get(c, V::first(vals), all_same);
    if (all_same) id = lid;
    // branch blocks: per-warp value-keyed cache, then the global table
    const u32 e = V::hash(eff) & (V::BC - 1);
    bool pend = need_branch;
    if (need_branch) {
        bool hit = true;
#pragma unroll
        for (int j = 0; j < V::KW; ++j) hit = hit && c.ws->bkey[e * V::KW + j] == eff.w[j];
        if (hit) {
            u64 v = c.ws->bval[e];
            if (v != 0) {
                id = v;
                pend = false;
                c.t.local++;
            }
        }
    }
    const u32 pmask = __ballot_sync(FULL, pend);
    if (pmask != 0) {
        // duplicate keys inside the warp: one representative (leader) per distinct value word
        u32 mm = 0;
        if (pend) {
            mm = __match_any_sync(pmask, eff.w[0]);
            if (V::KW > 1) {
#pragma unroll
                for (int j = 1; j < V::KW; ++j) mm &= __match_any_sync(pmask, eff.w[j]);
            }
        }
        const int leader_lane = pend ? (__ffs(mm) - 1) : 0;
        const bool leader = pend && leader_lane == lane;
        const u32 lmask = __ballot_sync(FULL, leader);
        u64 bid;
        if (c.tpk_only || __popc(lmask) > 4) {
            bid = intern_block<T>(c, leader, eff);  // many distinct misses: every leader probes on its own
        } else {
            // a few misses: one 8-lane group per key (group g takes the g-th leader)
            int src = __fns(lmask, 0, (c.gs >> 3) + 1);
            const bool gvalid = src >= 0 && src < 32;
            if (!gvalid) src = 0;
            typename V::Key gk = V::bcast(eff, src);
            const u32 cv = gvalid ? V::get(gk, c.li) : 0;
            u64 child = leaf_get(c, cv, gvalid);
            u64 gid = intern_branch<T>(c, gvalid, child, true, cv);
            // hand the id to the leader lane it belongs to
            bid = 0;
#pragma unroll
            for (int g2 = 0; g2 < 4; ++g2) {
                int s2 = __shfl_sync(FULL, src, g2 * 8);
                u64 id2 = __shfl_sync(FULL, gid, g2 * 8);
                bool v2 = __shfl_sync(FULL, int(gvalid), g2 * 8) != 0;
                if (v2 && lane == s2) bid = id2;
            }
        }
        // cache refresh: one writer per entry, so key and id of an entry always belong together
        const bool wr = leader && bid != 0;
        const u32 wb = __ballot_sync(FULL, wr);
        __syncwarp();  // the entry's earlier reads by other lanes are done (WAR)
        if (wr) {
            u32 sm = __match_any_sync(wb, e);
            if ((__ffs(sm) - 1) == lane) {
#pragma unroll
                for (int j = 0; j < V::KW; ++j) c.ws->bkey[e * V::KW + j] = eff.w[j];
                c.ws->bval[e] = bid;
            }
        }
        // every pending lane takes its leader's id
        u64 lid2 = __shfl_sync(FULL, bid, leader_lane);
        if (pend) id = lid2;
        __syncwarp();
    }
    return id;
}

// What stands at the unit before the batch: nothing, a fill leaf (phase 0), or an old tree.
struct Under {
    u64 old_root;  // OLD only
    u64 fill_leaf; // 0 = none
    u32 fill;
    int depth;     // tree depth D
};

// OLD: old node + old voxel values under each of the warp's 32 blocks (lane = block `blk`).
template <class T>
__device__ inline void old_block_state(Ctx<T>& c, const Under& u, u32 blk, bool active, bool wants_values,
                                       u64* parent_old, u64* old_id, typename VT<T>::Key* oldw) {
    using V = VT<T>;
    const InternerDev& in = c.in;
    // the level-1 parent's old node (group uniform: eight lanes walk the same path, loads coalesce)
    u64 E = old_at(in, u.old_root, blk >> 3, u.depth - 2);
    *parent_old = E;
    u64 ob = E;
    if (E != 0 && !id_is_leaf(E)) ob = ld_strong(&in.children[size_t(id_index(E)) * 8 + (blk & 7)]);
    if (!active) ob = 0;
    *old_id = ob;
    typename V::Key w = V::zero();
    const bool need = active && wants_values && ob != 0;
    if (need && id_is_leaf(ob)) w = V::splat(child_value<T>(in, ob));  // uniform under an old leaf
    // old block is a branch: its eight voxels are gathered by an 8-lane group, four blocks a round
    u32 bmask = __ballot_sync(FULL, need && !id_is_leaf(ob));
    while (bmask != 0) {
        int src = __fns(bmask, 0, (c.gs >> 3) + 1);
        const bool gvalid = src >= 0 && src < 32;
        if (!gvalid) src = 0;
        u64 gob = __shfl_sync(FULL, ob, src);
        u64 ch = gvalid ? ld_strong(&in.children[size_t(id_index(gob)) * 8 + c.li]) : 0;
        u32 v = (gvalid && ch != 0) ? child_value<T>(in, ch) : 0;
        typename V::Key gk = V::assemble(v, c.li);
        u32 done = 0;
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2) {
            typename V::Key k2 = V::bcast(gk, g2 * 8);
            int s2 = __shfl_sync(FULL, src, g2 * 8);
            bool v2 = __shfl_sync(FULL, int(gvalid), g2 * 8) != 0;
            if (v2) {
                done |= 1u << s2;
                if (c.lane == s2) w = k2;
            }
        }
        bmask &= ~done;
    }
    *oldw = w;
}

// ------------------------------------------------------------------------------------------------
// Block stage of a unit: `nblocks` (8, 64 or 512) Morton-consecutive blocks starting at block
// `first_block` of the chunk -> the unit's level-1 parents in ws->l1[0 .. nblocks/8) / ws->p1.
// Returns false when nothing in the unit has a set bit (fresh trees only).
// ------------------------------------------------------------------------------------------------
// All set_masks of a unit in one shot: lane L gets the masks of blocks [16L, 16L+16) — two 16-byte
// coalesced loads per lane, 1 KiB per warp for a full 512-block unit.  clear_mask bytes are dropped
// here: apply never reads them (voxtree.rs:778-781).
__device__ __forceinline__ void load_unit_masks(const u8* masks, size_t first_block, int nblocks, int lane, u64* mlo,
                                                u64* mhi) {
    *mlo = 0;
    *mhi = 0;
    const uint4* mp = (const uint4*)(masks + (first_block + 16 * lane) * 2);
    if (16 * lane + 8 <= nblocks) {
        uint4 q = ld_stream_v4(mp);
        *mlo = u64(__byte_perm(q.x, q.y, 0x6420)) | (u64(__byte_perm(q.z, q.w, 0x6420)) << 32);
    }
    if (16 * lane + 16 <= nblocks) {
        uint4 q = ld_stream_v4(mp + 1);
        *mhi = u64(__byte_perm(q.x, q.y, 0x6420)) | (u64(__byte_perm(q.z, q.w, 0x6420)) << 32);
    }
}

template <class T, bool OLD>
__device__ __forceinline__ bool build_blocks(Ctx<T>& c, u64 mlo, u64 mhi, const void* values, u32 first_block,
                                             int nblocks, const Under& u) {
    using V = VT<T>;
    const int lane = c.lane;
    const int n_iter = (nblocks + 31) / 32;
    const typename V::Key fillw = V::splat(u.fill_leaf ? u.fill : 0);
    // Which 32-block iterations contain any set bit (iteration it <- lanes 2it, 2it+1).  Blocks with
    // set_mask == 0 are skipped by phase 1 (voxtree.rs:779-781): their VALUES are never loaded, and
    // iterations without a set bit cost nothing.
    u32 nzl = __ballot_sync(FULL, (mlo | mhi) != 0);
    u32 iters = (nzl | (nzl >> 1)) & 0x55555555u;  // bit 2*it
    if (OLD) iters = (n_iter >= 16 ? 0xFFFFFFFFu : ((1u << (2 * n_iter)) - 1)) & 0x55555555u;  // every group needs its old node
    if (!OLD && iters == 0) return false;  // nothing set in this unit: it stays what it was (fill leaf / EMPTY)
    if (!OLD) {
        c.ws->l1[lane] = u.fill_leaf;
        c.ws->l1[lane + 32] = u.fill_leaf;
    }
    ((u16*)c.ws->p1)[lane] = 0;
    __syncwarp();
    auto mask_of = [&](int it) -> u32 {
        const int src = 2 * it + (lane >> 4);
        u64 lo = __shfl_sync(FULL, mlo, src), hi = __shfl_sync(FULL, mhi, src);
        u64 w = (lane & 8) ? hi : lo;
        return u32(w >> (8 * (lane & 7))) & 0xFF;
    };
    // software pipeline over the non-empty iterations: the value loads of the next one are issued
    // before the current one is processed; only lanes whose block has a set bit load (8 B / 32 B).
    int it = iters ? ((__ffs(iters) - 1) >> 1) : -1;
    typename V::Key nvals = V::zero();
    u32 nmask = 0;
    if (it >= 0) {
        nmask = mask_of(it);
        if (nmask) nvals = V::load(values, size_t(first_block) + it * 32 + lane);
    }
    while (it >= 0) {
        iters &= iters - 1;
        const int cur = it;
        typename V::Key vals = nvals;
        u32 smask = nmask;
        it = iters ? ((__ffs(iters) - 1) >> 1) : -1;
        if (it >= 0) {
            nmask = mask_of(it);
            nvals = V::zero();
            if (nmask) nvals = V::load(values, size_t(first_block) + it * 32 + lane);
        }
        const bool active = cur * 32 + lane < nblocks;
        typename V::Key oldw = fillw;
        u64 old_id = u.fill_leaf, parent_old = u.fill_leaf;
        if (OLD) old_block_state<T>(c, u, first_block + cur * 32 + lane, active, smask != 0, &parent_old, &old_id, &oldw);
        bool present;
        u64 id = block_node<T>(c, active, vals, smask, oldw, old_id, &present);
        bool pp;
        u64 pid = parent_node<T>(c, id, present && active, parent_old, &pp);
        if (c.li == 0 && cur * 32 + c.gs < nblocks) {
            c.ws->l1[cur * 4 + (c.gs >> 3)] = pid;
            c.ws->p1[cur * 4 + (c.gs >> 3)] = pp;
        }
    }
    __syncwarp();
    return true;
}

// True when all 512 blocks of unit `unit` have all eight voxels set (mlo / mhi: load_unit_masks) to ONE non-default
// value; lane L tests its 16 blocks = 128 * sizeof(T) contiguous bytes, all requested before the first compare.
template <class T>
__device__ __forceinline__ bool solid_unit(const void* chunk_values, u32 unit, u64 mlo, u64 mhi, int lane, u32* value) {
    if (!__all_sync(FULL, (mlo & mhi) == ~0ull)) return false;
    const uint4* vp = reinterpret_cast<const uint4*>((const u8*)chunk_values + (size_t(unit) * UNIT_BLOCKS + 16 * lane) * 8 * sizeof(T));
    constexpr int NV = 8 * int(sizeof(T));  // 16-byte vectors holding this lane's 16 blocks
    uint4 q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = ld_stream_v4(vp + j);
    const u32 v0 = __shfl_sync(FULL, sizeof(T) == 1 ? (q[0].x & 0xFFu) : q[0].x, 0);
    const u32 splat = sizeof(T) == 1 ? v0 * 0x01010101u : v0;
    bool uni = v0 != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) uni = uni && q[j].x == splat && q[j].y == splat && q[j].z == splat && q[j].w == splat;
#pragma unroll 1
    for (int j0 = 8; j0 < NV && __all_sync(FULL, uni); j0 += 8) {  // wider T: the rest of the lane's 16 blocks
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = ld_stream_v4(vp + j0 + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) uni = uni && q[j].x == splat && q[j].y == splat && q[j].z == splat && q[j].w == splat;
    }
    *value = v0;
    return __all_sync(FULL, uni);
}

// Joins ws->l1[0..n) (nodes of depth d; node i sits at position pos0+i of that depth) level by level
// down to one node, compacting in place.  The ONLY other call site of parent_node (keeps the kernel
// inside the instruction cache).
template <class T>
__device__ __forceinline__ u64 reduce_levels(Ctx<T>& c, u32 n, int d, u32 pos0, const Under& u, bool use_old,
                                             bool* present) {
    while (n > 1) {
        for (u32 it = 0; it * 32 < n; ++it) {
            const u32 i = it * 32 + c.lane;
            const bool act = i < n;
            u64 ch = act ? c.ws->l1[i] : 0;
            bool pr = act && c.ws->p1[i] != 0;
            u64 old_self = u.fill_leaf;
            if (use_old) old_self = old_at(c.in, u.old_root, (pos0 + it * 32 + c.gs) >> 3, d - 1);
            bool pp;
            u64 pid = parent_node<T>(c, ch, pr, old_self, &pp);
            __syncwarp();  // all reads of this slice are done before its head is overwritten
            if (c.li == 0 && it * 32 + c.gs < n) {
                c.ws->l1[it * 4 + (c.gs >> 3)] = pid;
                c.ws->p1[it * 4 + (c.gs >> 3)] = pp;
            }
        }
        __syncwarp();
        n >>= 3;
        d -= 1;
        pos0 >>= 3;
    }
    *present = c.ws->p1[0] != 0;
    u64 r = c.ws->l1[0];
    __syncwarp();
    return r;
}

// ------------------------------------------------------------------------------------------------
// Kernel plumbing
// ------------------------------------------------------------------------------------------------
struct ApplyArgs {
    InternerDev in;
    const u8* masks;         // [n][B][2]
    const void* values;      // [n][B][8]
    const u8* flags;         // [n] or null
    const long long* fills;  // [n] or null
    const u64* old_roots;    // [n] or null (OLD kernels only): root of the tree before the batch
    u64* roots;              // [n]
    u8* changed;             // [n] or null
    // join scratch (D >= 5): units of 512 blocks are built by independent warps; the last warp to
    // finish a 32^3 cube joins its eight unit nodes, the last cube of a chunk joins the top levels
    u64* unit_ids;      // [n][units_per_chunk]
    u8* unit_present;   // [n][units_per_chunk]
    u32* cube_done;     // [n][cubes_per_chunk]  zeroed before launch
    u64* cube_ids;      // [n][cubes_per_chunk]
    u8* cube_present;   // [n][cubes_per_chunk]
    u32* chunk_done;    // [n]                   zeroed before launch
    u32* work_next;     // dynamic work counter (runs), zeroed before launch
    u32 run;            // units per work item: 8 = a whole 32^3 cube per warp (local join), 1 = one unit
    u32 n;
    u32 depth;
    u32 blocks;    // B
    u32 use_free;  // free list non-empty at launch
};

template <class T>
__device__ __forceinline__ void ctx_init(Ctx<T>& c, const InternerDev& in, WarpSmem<T>* ws, CtaSmem* cs, bool use_free) {
    c.in = in;
    c.use_free = use_free;
    c.tpk_only = false;
    c.weak_first = false;
    c.lane = threadIdx.x & 31;
    c.li = c.lane & 7;
    c.gs = c.lane & 24;
    c.ws = ws + (threadIdx.x >> 5);
    c.cs = cs;
}

// block_level = false: the kernel never looks a block up by its value word (level / upper kernels of the bulk builder),
// so the per-warp block cache — 60 % of the shared memory — is left as it is.
template <class T>
__device__ inline void smem_init(WarpSmem<T>* ws, CtaSmem* cs, bool block_level = true) {
    if (block_level) {
        u32* w = (u32*)ws;
        for (u32 i = threadIdx.x; i < sizeof(WarpSmem<T>) * WARPS_PER_CTA / 4; i += blockDim.x) w[i] = 0;
    } else {
        constexpr u32 words = u32((sizeof(WarpSmem<T>) - offsetof(WarpSmem<T>, ukey)) / 4);
        u32* tail = (u32*)((unsigned char*)(ws + (threadIdx.x >> 5)) + offsetof(WarpSmem<T>, ukey));  // every warp its own
        for (u32 i = threadIdx.x & 31; i < words; i += 32) tail[i] = 0;
    }
    u32* q = (u32*)cs;
    for (u32 i = threadIdx.x; i < sizeof(CtaSmem) / 4; i += blockDim.x) q[i] = 0;
    __syncthreads();
}

// Kernel exit.  No CTA barrier: warps leave as they run out of work; the last one to arrive flushes
// the CTA's leaf in-degree histogram.
template <class T>
__device__ inline void cta_finish(Ctx<T>& c) {
    Tally& t = c.t;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        t.leaf_calls += __shfl_xor_sync(FULL, t.leaf_calls, o);
        t.branch_calls += __shfl_xor_sync(FULL, t.branch_calls, o);
        t.leaf_miss += __shfl_xor_sync(FULL, t.leaf_miss, o);
        t.branch_miss += __shfl_xor_sync(FULL, t.branch_miss, o);
        t.collapsed += __shfl_xor_sync(FULL, t.collapsed, o);
        t.probes += __shfl_xor_sync(FULL, t.probes, o);
        t.local += __shfl_xor_sync(FULL, t.local, o);
    }
    if (c.lane == 0) {
        Counters* k = c.in.ctr;
        if (t.leaf_calls) atomicAdd(&k->leaf_calls, (ull)t.leaf_calls);
        if (t.branch_calls) atomicAdd(&k->branch_calls, (ull)t.branch_calls);
        if (t.leaf_miss) atomicAdd(&k->leaf_misses, (ull)t.leaf_miss);
        if (t.branch_miss) atomicAdd(&k->branch_misses, (ull)t.branch_miss);
        if (t.collapsed) atomicAdd(&k->collapsed, (ull)t.collapsed);
        if (t.probes) atomicAdd(&k->probe_steps, (ull)t.probes);
        if (t.local) atomicAdd(&k->cache_hits_local, (ull)t.local);
    }
    {
        u32 arrived = 0;
        __syncwarp();
        if (c.lane == 0) {
            __threadfence_block();
            arrived = atomicAdd(&c.cs->warps_done, 1u);
        }
        arrived = __shfl_sync(FULL, arrived, 0);
        if (arrived == WARPS_PER_CTA - 1) {
            __threadfence_block();
            if (sizeof(T) == 1) {
                for (u32 v = c.lane; v < 256; v += 32) {
                    u32 n = ((volatile u32*)c.cs->leafref)[v];
                    if (n) atomicAdd(&c.in.refs[id_index(lds_relaxed(&c.cs->leaf[v]))], n);
                }
            }
            for (u32 i = c.lane; i < RH; i += 32) {
                u32 k = ((volatile u32*)c.cs->rkey)[i], n = ((volatile u32*)c.cs->rcnt)[i];
                if (k != 0 && n != 0) atomicAdd(&c.in.refs[k - 1], n);
            }
        }
    }
}

// Phase 0 (voxtree.rs:742-758) for one chunk: returns the fill leaf (0 if none) to the whole warp.
// The reference's get_or_create_leaf(fill) hands out one reference that is either the root handle
// (no patches) or never released — reproduced by `take_ref`.
template <class T>
__device__ inline u64 phase0_fill(Ctx<T>& c, bool has_fill, u32 fill, bool take_ref) {
    u64 fl = leaf_get(c, fill, has_fill && c.lane == 0);
    fl = __shfl_sync(FULL, fl, 0);
    if (has_fill && take_ref && c.lane == 0 && fl != 0) {
        atomicAdd(&c.in.refs[id_index(fl)], 1u);
        c.t.leaf_calls++;
    }
    return fl;
}

template <class T>
__device__ __forceinline__ void chunk_flags(const ApplyArgs& a, u32 chunk, bool* has_fill, u32* fill, bool* has_patches) {
    u32 f = a.flags ? a.flags[chunk] : VX_FLAG_PATCHES;
    *has_fill = (f & VX_FLAG_FILL) != 0;
    *has_patches = (f & VX_FLAG_PATCHES) != 0;
    long long fv = (*has_fill && a.fills) ? a.fills[chunk] : 0;
    *fill = sizeof(T) == 1 ? u32(fv) & 0xFF : u32(fv);
    if (*fill == 0) *has_fill = false;  // fill(default) == clear: nothing to intern
}

template <class T>
__device__ __forceinline__ void write_root(Ctx<T>& c, const ApplyArgs& a, u32 chunk, u64 root, bool changed) {
    // apply_batch (voxtree.rs:303-328): INVALID -> false and the tree keeps its root.  The caller
    // owns the old root's release (dec_ref_recursive, :309-314).
    a.roots[chunk] = changed ? root : 0;
    if (a.changed) a.changed[chunk] = changed ? 1 : 0;
    if (changed && root != 0) atomicAdd(&c.in.refs[id_index(root)], 1u);  // the tree's root handle
}

// ------------------------------------------------------------------------------------------------
// apply kernel.  Unit = <= 512 Morton-consecutive blocks (a 16^3-voxel sub-cube, or the whole chunk
// when D <= 4), built by ONE warp.  Work items are runs of `a.run` consecutive units handed out by an
// atomic counter (dynamic load balance: units range from empty to 512 new nodes):
//   run == 8  a warp owns a whole 32^3 cube and joins its eight unit nodes locally;
//   run == 1  (few chunks) units of a cube go to different warps; each publishes its node to global
//             scratch and the LAST to arrive (atomic counter) joins the cube.
// The last cube of a chunk joins the top levels the same way (D = 6: 8 cubes, D = 7: 64).
// No __syncthreads anywhere in the loop: warps never wait for each other.
// ------------------------------------------------------------------------------------------------
#ifndef VX_MIN_CTAS
#define VX_MIN_CTAS 3
#endif
template <class T>
constexpr size_t apply_smem_bytes() {
    return sizeof(WarpSmem<T>) * WARPS_PER_CTA + sizeof(CtaSmem);
}
template <class T, bool OLD>
__global__ void __launch_bounds__(CTA_THREADS, VX_MIN_CTAS) apply_kernel(ApplyArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];  // may exceed 48 KiB: dynamic (apply_smem_bytes)
    WarpSmem<T>* ws = reinterpret_cast<WarpSmem<T>*>(smem_raw);
    CtaSmem* csp = reinterpret_cast<CtaSmem*>(smem_raw + sizeof(WarpSmem<T>) * WARPS_PER_CTA);
    smem_init<T>(ws, csp);
    Ctx<T> c;
    ctx_init<T>(c, a.in, ws, csp, a.use_free != 0);
    const int lane = c.lane;
    const int D = int(a.depth);
    const u32 upc = a.blocks > UNIT_BLOCKS ? a.blocks / UNIT_BLOCKS : 1;  // units per chunk: 1, 8, 64, 512
    const int upc_log = 31 - __clz(upc);
    const int nblocks = a.blocks > UNIT_BLOCKS ? UNIT_BLOCKS : int(a.blocks);
    const u32 cpc = upc / 8;                                             // 32^3 cubes per chunk: 0, 1, 8, 64
    const u32 R = a.run;
    const unsigned long long total_units = (unsigned long long)a.n * upc;
    const unsigned long long total_runs = (total_units + R - 1) / R;

    auto grab = [&]() -> unsigned long long {  // next run index (dynamic)
        u32 r = 0;
        if (lane == 0) r = atomicAdd(a.work_next, 1u);
        return __shfl_sync(FULL, r, 0);
    };
    auto masks_of = [&](unsigned long long w, u64* lo, u64* hi) {
        load_unit_masks(a.masks + size_t(w >> upc_log) * a.blocks * 2, size_t(w & (upc - 1)) * UNIT_BLOCKS, nblocks, lane, lo,
                        hi);
    };

    unsigned long long run = grab();
    unsigned long long next_run = run < total_runs ? grab() : total_runs;
    u64 nlo = 0, nhi = 0;
    if (R != 8 && run < total_runs) masks_of(run * R, &nlo, &nhi);
    while (run < total_runs) {
        const bool poisoned = __any_sync(FULL, lane == 0 && ld_strong(a.in.error) != ERR_NONE);
        // run == 8: one pass over the cube's 8 KiB of masks (already pulled into L2) tells which of its
        // eight units have a set bit at all; units without one cost nothing more than this test
        u32 ne_units = 0xFF;
        if (R == 8 && !poisoned) {
            const unsigned long long w0 = run * 8;
            const uint4* mp = (const uint4*)(a.masks + (size_t(w0 >> upc_log) * a.blocks + size_t(u32(w0) & (upc - 1)) * UNIT_BLOCKS) * 2) + lane * 2;
            ne_units = 0;
#pragma unroll 2
            for (int k2 = 0; k2 < 8; ++k2) {
                const uint4 q0 = ld_stream_v4(mp + k2 * 64), q1 = ld_stream_v4(mp + k2 * 64 + 1);
                const bool any = ((q0.x | q0.y | q0.z | q0.w | q1.x | q1.y | q1.z | q1.w) & 0x00FF00FFu) != 0;  // set_mask bytes only
                ne_units |= u32(__any_sync(FULL, any)) << k2;
            }
        }
        for (u32 k = 0; k < R; ++k) {
            const unsigned long long w = run * R + k;
            if (w >= total_units) break;
            const u32 chunk = u32(w >> upc_log), unit = u32(w) & (upc - 1);
            u64 mlo = 0, mhi = 0;
            if (R == 8) {
                if ((ne_units >> k) & 1) masks_of(w, &mlo, &mhi);  // L1/L2 hit: the pre-scan just read them
            } else {
                mlo = nlo;
                mhi = nhi;
                // request the masks of the unit after this one before doing any work
                if (next_run < total_runs) masks_of(next_run * R, &nlo, &nhi);
            }
            // poisoned interner: skip the rest, the host reports the error (warp-uniform, checked per run)
            if (poisoned) {
                if (lane == 0 && unit == 0) {
                    a.roots[chunk] = 0;
                    if (a.changed) a.changed[chunk] = 0;
                }
                continue;
            }
            bool has_fill, has_patches;
            u32 fill;
            chunk_flags<T>(a, chunk, &has_fill, &fill, &has_patches);
            // phase 0 once per chunk (unit 0 takes the reference); every unit needs the fill leaf's id
            u64 fl = phase0_fill<T>(c, has_fill, fill, has_patches && unit == 0);
            if (!has_patches) {
                // :756-758 returns the initial node: Leaf(fill), or — no fill either — the tree's own root.  On
                // an EMPTY tree that root is EMPTY, which is not INVALID, so apply_batch (:303-328) reports
                // "changed" and marks the tree dirty with nothing built.  On a non-empty tree the reference
                // trips its assert_ne!(new_root, old_root) (:310); here that case is "unchanged".
                if (lane == 0 && unit == 0) {
                    if (has_fill) c.t.leaf_calls++;
                    const bool old_empty = !(OLD && a.old_roots && a.old_roots[chunk] != 0);
                    write_root<T>(c, a, chunk, fl, has_fill || old_empty);
                }
                continue;
            }
            Under u{0, fl, fill, D};
            // with a fill the batch is built against Leaf(fill); the old tree is only released (:742-754)
            const bool use_old = OLD && !has_fill && a.old_roots && a.old_roots[chunk] != 0;
            if (use_old) u.old_root = a.old_roots[chunk];
            const void* cv = (const u8*)a.values + size_t(chunk) * a.blocks * 8 * sizeof(T);

            if (R == 8 && ne_units == 0 && !use_old && cpc == 1) {
                // nothing set anywhere in this 32^3 chunk: no block enters `paths` (voxtree.rs:905-911)
                if (lane == 0) write_root<T>(c, a, chunk, fl, false);
                break;
            }
            // ---- stage 0: the unit.  stage 1: its 32^3 cube.  stage 2: the chunk's top levels.
            bool some = false;
            u64 solid = 0;
            if (use_old)
                some = build_blocks<T, OLD>(c, mlo, mhi, cv, unit * UNIT_BLOCKS, nblocks, u);
            else if (R != 8 || ((ne_units >> k) & 1)) {
                // Solid unit (set_uniform, filled volumes): every voxel set to the same non-default value.  Phase 1 makes
                // 512 identical leaves (:826 does not look at what stood there), phase 2 collapses 64 + 8 + 1 parents
                // (:1050) — the outcome is that one leaf, and the counters are those of the long way round.
                u32 v0 = 0;
                if (nblocks == UNIT_BLOCKS && solid_unit<T>(cv, unit, mlo, mhi, lane, &v0)) {
                    solid = __shfl_sync(FULL, leaf_get(c, v0, lane == 0), 0);
                    if (lane == 0) {
                        c.t.leaf_calls += UNIT_BLOCKS;
                        c.t.collapsed += UNIT_BLOCKS + UNIT_BLOCKS / 8 + UNIT_BLOCKS / 64 + 1;
                    }
                } else {
                    some = build_blocks<T, false>(c, mlo, mhi, cv, unit * UNIT_BLOCKS, nblocks, u);
                }
            }
            u32 rn = u32(nblocks) / 8, rpos = (unit * UNIT_BLOCKS) >> 3;
            int rd = D - 2;
            for (int stage = 0;; ++stage) {
                u64 node = u.fill_leaf;
                bool present = false;
                if (stage == 0 && solid != 0) {
                    node = solid;
                    present = true;
                } else if (some)
                    node = reduce_levels<T>(c, rn, rd, rpos, u, use_old, &present);
                if (stage == 0) {
                    if (upc == 1) {  // D <= 4: the unit is the tree
                        if (lane == 0) write_root<T>(c, a, chunk, node, present);
                        break;
                    }
                    const u32 cube = unit >> 3;
                    if (R == 8) {  // this warp owns the whole cube: keep the unit nodes in shared memory
                        if (lane == 0) {
                            c.ws->run_ids[unit & 7] = node;
                            c.ws->run_pres[unit & 7] = present;
                        }
                        __syncwarp();
                        if ((unit & 7) != 7) break;
                        if (lane < 8) {
                            c.ws->l1[lane] = c.ws->run_ids[lane];
                            c.ws->p1[lane] = c.ws->run_pres[lane];
                        }
                    } else {  // publish the unit; the last arriver of the cube joins it
                        u32 arrived = 0;
                        if (lane == 0) {
                            st_strong(&a.unit_ids[w], node);
                            ((volatile u8*)a.unit_present)[w] = present;
                            fence_gpu();
                            arrived = atomicAdd(&a.cube_done[size_t(chunk) * cpc + cube], 1u);
                        }
                        arrived = __shfl_sync(FULL, arrived, 0);
                        if (arrived != 7) break;
                        fence_gpu();
                        const size_t base = (size_t(chunk) * cpc + cube) * 8;
                        if (lane < 8) {
                            c.ws->l1[lane] = ld_strong(&a.unit_ids[base + lane]);
                            c.ws->p1[lane] = ((volatile u8*)a.unit_present)[base + lane];
                        }
                    }
                    __syncwarp();
                    some = true;  // eight 16^3 units (depth D-4) -> the cube's node (depth D-5)
                    rn = 8;
                    rd = D - 4;
                    rpos = cube * 8;
                } else if (stage == 1) {
                    if (cpc == 1) {
                        if (lane == 0) write_root<T>(c, a, chunk, node, present);
                        break;
                    }
                    // D >= 6: publish the cube; the last cube of the chunk joins the top levels
                    const u32 cube = unit >> 3;
                    u32 arrived = 0;
                    if (lane == 0) {
                        st_strong(&a.cube_ids[size_t(chunk) * cpc + cube], node);
                        ((volatile u8*)a.cube_present)[size_t(chunk) * cpc + cube] = present;
                        fence_gpu();
                        arrived = atomicAdd(&a.chunk_done[chunk], 1u);
                    }
                    arrived = __shfl_sync(FULL, arrived, 0);
                    if (arrived != cpc - 1) break;
                    fence_gpu();
                    for (u32 i = lane; i < cpc; i += 32) {
                        c.ws->l1[i] = ld_strong(&a.cube_ids[size_t(chunk) * cpc + i]);
                        c.ws->p1[i] = ((volatile u8*)a.cube_present)[size_t(chunk) * cpc + i];
                    }
                    __syncwarp();
                    some = true;  // cube nodes (depth D-5) -> root
                    rn = cpc;
                    rd = D - 5;
                    rpos = 0;
                } else {
                    if (lane == 0) write_root<T>(c, a, chunk, node, present);
                    break;
                }
            }
        }
        run = next_run;
        next_run = run < total_runs ? grab() : total_runs;
        if (next_run < total_runs) {
            // pull the masks of the run after the next into L2 (R units x 2*nblocks bytes, 128-byte lines):
            // empty units are so cheap that one unit of register prefetch does not cover DRAM latency
            const u8* mp = a.masks + (size_t((next_run * R) >> upc_log) * a.blocks + size_t((next_run * R) & (upc - 1)) * UNIT_BLOCKS) * 2;
            const u32 bytes = R * u32(nblocks) * 2;
            for (u32 off = u32(lane) * 128; off < bytes; off += 32 * 128) prefetch_l2(mp + off);
        }
    }
    cta_finish<T>(c);
}

}  // namespace vx

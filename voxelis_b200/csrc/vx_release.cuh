// vx_release.cuh — releasing trees: dec_ref_recursive / recycle on the device.
//
// Replaces  VoxInterner::dec_ref_recursive  interner/mod.rs:419-534  (FIFO work list from the old root)
//           VoxInterner::recycle            interner/mod.rs:566-625  (zero the slot, generation++, wrap at
//                                                                    MAX_GENERATION 0x7FFE, push the free list)
//           pattern-map removal             interner/mod.rs:276-281,405-412,520-526
// The reference walks one tree serially; here a level-synchronous frontier does it for any number of roots:
// a node enters the frontier exactly when an atomic decrement takes its refcount to zero, so every node
// is freed once no matter how many parents release it in the same round.  One 8-lane group frees one
// node: the children row is read and cleared coalesced, each lane drops the reference on its child.
#pragma once
#include "vx_device.cuh"

namespace vx {

constexpr u16 MAX_GENERATION = 0x7FFE;  // core/block_id.rs:130

// The tree handles give up their reference; roots that die start the frontier.
__global__ void release_roots_kernel(InternerDev in, const u64* roots, u32 n, u64* frontier, u32* frontier_count) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 id = roots[i];
    if (id == 0) return;
    u32 old = atomicSub(&in.refs[id_index(id)], 1u);
    if (old == 1u) frontier[atomicAdd(frontier_count, 1u)] = id;
}

// tree.set_root_id (voxtree.rs:135-141): one more handle on `id`.
__global__ void add_ref_kernel(InternerDev in, u64 id, u32 by) {
    if (id != 0) atomicAdd(&in.refs[id_index(id)], by);
}

// After an apply that popped the free list concurrently the counter may have gone below zero.
__global__ void clamp_free_count_kernel(InternerDev in) {
    if (int(*in.free_count) < 0) *in.free_count = 0;
}

template <class T>
__global__ void release_level_kernel(InternerDev in, const u64* frontier, u32 count, u64* next, u32* next_count) {
    const int lane = threadIdx.x & 31, li = lane & 7;
    const u32 group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const u32 ngroups = (gridDim.x * blockDim.x) >> 3;
    for (u32 base = 0; base < count; base += ngroups) {  // warp-uniform trip count
        const u32 gi = base + group;
        const bool valid = gi < count;
        const u64 id = valid ? frontier[gi] : 0;
        const u32 idx = id_index(id);
        if (valid && !id_is_leaf(id)) {
            u64 child = ld_strong(&in.children[size_t(idx) * 8 + li]);
            in.children[size_t(idx) * 8 + li] = 0;
            if (child != 0) {
                u32 old = atomicSub(&in.refs[id_index(child)], 1u);
                if (old == 1u) next[atomicAdd(next_count, 1u)] = child;
            }
        }
        if (valid && li == 0) {
            if (id_is_leaf(id)) {
                if (sizeof(T) == 1) {
                    u32 v = ld_strong_u8((const u8*)in.values + idx);
                    st_strong(&in.leaf_u8[v], 0);
                } else {
                    u32 v = ld_strong((const u32*)in.values + idx);
                    u64 key = u64(v) | (1ull << 32);
                    u32 s = u32(leaf_hash(v)) & in.leaf_mask;
                    for (u32 guard = 0; guard <= in.leaf_mask; ++guard) {
                        u64 k = ld_strong(&in.leaf_keys[s]);
                        if (k == key) {
                            st_strong(&in.leaf_ids[s], 0);
                            st_strong(&in.leaf_keys[s], 2ull << 32);  // tombstone: never equals a key
                            break;
                        }
                        if (k == 0) break;
                        s = (s + 1) & in.leaf_mask;
                    }
                }
            } else {
                u64 h = in.hashes[idx];
                u32 bucket = u32(h) & in.bucket_mask;
                bool found = false;
                for (u32 guard = 0; guard <= in.bucket_mask && !found; ++guard) {
                    bool any_empty = false;
                    for (int k = 0; k < 8; ++k) {
                        u64 slot = ld_strong(&in.slots[size_t(bucket) * 8 + k]);
                        if (slot == 0) any_empty = true;
                        if (u32(slot) == idx && slot != 0) {
                            st_strong(&in.slots[size_t(bucket) * 8 + k], u64(IDX_TOMB));
                            found = true;
                            break;
                        }
                    }
                    if (any_empty) break;
                    bucket = (bucket + 1) & in.bucket_mask;
                }
            }
            // recycle (mod.rs:566-625)
            if (sizeof(T) == 1)
                ((u8*)in.values)[idx] = 0;
            else
                ((u32*)in.values)[idx] = 0;
            in.hashes[idx] = 0;
            in.refs[idx] = 0;
            u16 g = in.gens[idx] + 1;
            if (g >= MAX_GENERATION) g = 0;
            in.gens[idx] = g;
            u32 pos = atomicAdd(in.free_count, 1u);
            in.free_list[pos] = idx;
            atomicAdd(&in.ctr->recycled, 1ull);
        }
    }
}

// Re-inserts every live node into cleared tables (drops accumulated tombstones): branches into the bucketised branch
// table, wide-T leaves into the open-addressing leaf table (released leaves leave tombstones there too — the
// reference's pattern map simply removes the entry, interner/mod.rs:276-281).  u8 leaves sit in a direct-mapped table.
template <class T>
__global__ void rehash_kernel(InternerDev in, u32 next_index) {
    u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 || idx >= next_index) return;
    u64 h = in.hashes[idx];
    if (h == 0) return;  // free slot
    bool branch = false;
    for (int k = 0; k < 8; ++k) branch = branch || in.children[size_t(idx) * 8 + k] != 0;
    if (!branch) {
        if (sizeof(T) == 1) return;
        const u32 v = ((const u32*)in.values)[idx];
        const u64 key = u64(v) | (1ull << 32);
        u32 s = u32(leaf_hash(v)) & in.leaf_mask;
        for (;;) {
            if (atomicCAS((ull*)&in.leaf_keys[s], 0ull, (ull)key) == 0ull) {
                in.leaf_ids[s] = id_leaf((u64(in.gens[idx]) << 32) | idx);
                return;
            }
            s = (s + 1) & in.leaf_mask;
        }
    }
    u64 word = (u64(u32(h >> 47)) << 47) | (u64(in.gens[idx]) << 32) | idx;
    u32 bucket = u32(h) & in.bucket_mask;
    for (;;) {
        for (int k = 0; k < 8; ++k)
            if (atomicCAS((ull*)&in.slots[size_t(bucket) * 8 + k], 0ull, (ull)word) == 0ull) return;
        bucket = (bucket + 1) & in.bucket_mask;
    }
}


// ------------------------------------------------------------------------------------------------
// vx_interner_reset without touching what was never used.  A fresh-build loop (build a world, reset, build the next)
// fills a small part of a table sized for the budget; clearing the whole table and the refcount pool costs 87 MB of
// stores per reset of a 256 MiB interner, 33 us next to a 0.26 ms build.  When no node has ever been released (no
// tombstones, no free list, generations all 0 — the host knows) the state to undo is exactly: the table slot of every
// branch in [1, next_index), the leaf-table entry of every leaf, and the refcounts.  One thread per node finds its own
// slot by its stored hash (it must exist, so the scan does not stop at empties other threads have just made).
// Falls back to clearing everything, in this same launch, when the interner is more than a few per cent full or its
// error word is set (an aborted insert may have left a claimed key behind).
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) reset_used_kernel(InternerDev in, size_t nbuckets, size_t leaf_slots) {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nt = size_t(gridDim.x) * blockDim.x;
    const u32 next = min(*in.next_index, in.capacity);
    const bool everything = *in.error != ERR_NONE || size_t(next) * 170 > nbuckets * 64 + size_t(in.capacity) * 6;
    if (sizeof(T) == 1)
        for (size_t i = t; i < 256; i += nt) in.leaf_u8[i] = 0;
    if (everything) {
        ulonglong2* sl = reinterpret_cast<ulonglong2*>(in.slots);
        for (size_t i = t; i < nbuckets * 4; i += nt) sl[i] = make_ulonglong2(0, 0);
        for (size_t i = t; i < in.capacity; i += nt) {
            in.refs[i] = 0;
            in.gens[i] = 0;
        }
        if (sizeof(T) != 1)
            for (size_t i = t; i < leaf_slots; i += nt) {
                in.leaf_keys[i] = 0;
                in.leaf_ids[i] = 0;
            }
        return;
    }
    for (size_t idx = t + 1; idx < next; idx += nt) {
        const u64 h = in.hashes[idx];
        in.refs[idx] = 0;
        if (h == 0) continue;
        bool branch = false;
        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(in.children + idx * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const ulonglong2 q = row[k];
            branch = branch || (q.x | q.y) != 0;
        }
        if (branch) {
            ulonglong2* wrow = reinterpret_cast<ulonglong2*>(in.children + idx * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) wrow[k] = make_ulonglong2(0, 0);  // a dead index keeps no old key (see ld_weak, vx_device.cuh)
            u32 bucket = u32(h) & in.bucket_mask;
            for (u32 guard = 0; guard <= in.bucket_mask; ++guard) {
                bool found = false;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const u64 slot = ld_strong(&in.slots[size_t(bucket) * 8 + k]);
                    if (slot != 0 && u32(slot) == u32(idx)) {
                        st_strong(&in.slots[size_t(bucket) * 8 + k], 0);
                        found = true;
                    }
                }
                if (found) break;
                bucket = (bucket + 1) & in.bucket_mask;
            }
        } else if (sizeof(T) != 1) {
            const u32 v = ((const u32*)in.values)[idx];
            const u64 key = u64(v) | (1ull << 32);
            u32 s = u32(leaf_hash(v)) & in.leaf_mask;
            for (u32 guard = 0; guard <= in.leaf_mask; ++guard) {
                if (ld_strong(&in.leaf_keys[s]) == key) {
                    st_strong(&in.leaf_ids[s], 0);
                    st_strong(&in.leaf_keys[s], 0);
                    break;
                }
                s = (s + 1) & in.leaf_mask;
            }
        }
    }
}

}  // namespace vx

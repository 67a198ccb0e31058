// vx_dedup.cuh — global dedup across per-GPU interners (BASELINE.json config 5).
//
// The reference has one interner for the whole world (world/voxmodel.rs:31-32) and its authors name
// the single pattern map as the scaling limit, proposing to shard it by `hash % N` (Voxelis Bible §3.9,
// §13).  This is that proposal on GPUs: after the per-GPU builds, every node is sent to the GPU that
// OWNS its key, `owner = hash(children as GLOBAL ids) mod G`, and interned there in a "global shard".
// Rounds are height-synchronous (leaves first), because a parent's key is made of its children's
// global ids: round h packs the local nodes of height h, an all-to-all (NCCL, done by the caller)
// moves 72-byte records to their owners, the owners intern them and answer with 8-byte global ids.
//
//   global id = the owner shard's BlockId with the owner's rank in generation bits [44..46]
//               (generations are 0 in a freshly merged shard, so the field is free).
//
// Kernels here: node heights, count / fill of the per-owner send buckets, owner-side interning of
// records (thread-per-key, same bucketised table and publication protocol as vx_build.cuh).
#pragma once
#include "vx_device.cuh"

namespace vx {

constexpr int DEDUP_REC_WORDS = 9;  // 8 child global ids + value
constexpr u32 H_DEAD = 255;         // free slot / not yet known

__host__ __device__ inline u64 with_owner(u64 id, u32 rank) { return id == 0 ? 0 : (id | (u64(rank & 7) << 44)); }
__host__ __device__ inline u32 owner_of_id(u64 gid) { return u32(gid >> 44) & 7; }

// heights: 0 = leaf, h >= 1 = branch whose tallest child has height h-1.  One sweep resolves every
// branch whose children are resolved; the host runs sweeps until nothing changes (<= depth + 1).
template <class T>
__global__ void heights_init_kernel(InternerDev in, u32 n, u8* h) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || in.hashes[i] == 0) {
        h[i] = H_DEAD;  // slot 0 (the empty branch) and recycled slots take no part
        return;
    }
    bool branch = false;
    for (int k = 0; k < 8; ++k) branch = branch || in.children[size_t(i) * 8 + k] != 0;
    h[i] = branch ? H_DEAD : 0;
}
__global__ void heights_sweep_kernel(InternerDev in, u32 n, u8* h, u32* changed) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0 || h[i] != H_DEAD || in.hashes[i] == 0) return;
    u32 mx = 0;
    for (int k = 0; k < 8; ++k) {
        u64 ch = in.children[size_t(i) * 8 + k];
        if (ch == 0) continue;
        u32 hc = ((volatile u8*)h)[id_index(ch)];
        if (hc == H_DEAD) return;  // a child is not resolved yet: next sweep
        mx = max(mx, hc + 1);
    }
    h[i] = u8(mx);
    *changed = 1;
}

// key of local node i in global ids -> (hash, owner)
template <class T>
__device__ __forceinline__ u64 global_key_hash(const InternerDev& in, u32 i, const u64* gmap, bool leaf, u64* kids) {
    if (leaf) {
        u64 v = sizeof(T) == 1 ? u64(((const u8*)in.values)[i]) : u64(((const u32*)in.values)[i]);
        return leaf_hash(v);
    }
    u64 h = 0;
    for (int k = 0; k < 8; ++k) {
        u64 ch = in.children[size_t(i) * 8 + k];
        u64 g = ch ? gmap[id_index(ch)] : 0;
        kids[k] = g;
        h += child_hash(g, k);
    }
    return finish_hash(h);
}

// pass 1: how many nodes of height `height` go to each owner
template <class T>
__global__ void dedup_count_kernel(InternerDev in, u32 n, const u8* h, u32 height, const u64* gmap, u32 G, u32* counts) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || h[i] != height) return;
    u64 kids[8];
    u64 hash = global_key_hash<T>(in, i, gmap, height == 0, kids);
    atomicAdd(&counts[u32(hash >> 40) % G], 1u);
}
// pass 2: write the 72-byte records grouped by owner (bases = exclusive scan of counts)
template <class T>
__global__ void dedup_fill_kernel(InternerDev in, u32 n, const u8* h, u32 height, const u64* gmap, u32 G,
                                  const u32* bases, u32* cursors, u64* records, u32* src) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || h[i] != height) return;
    u64 kids[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u64 hash = global_key_hash<T>(in, i, gmap, height == 0, kids);
    u32 o = u32(hash >> 40) % G;
    u32 pos = bases[o] + atomicAdd(&cursors[o], 1u);
    u64* r = records + size_t(pos) * DEDUP_REC_WORDS;
    for (int k = 0; k < 8; ++k) r[k] = kids[k];
    r[8] = sizeof(T) == 1 ? u64(((const u8*)in.values)[i]) : u64(((const u32*)in.values)[i]);
    src[pos] = i;
}
// after the answers came back: gmap[src[j]] = ids[j]
__global__ void dedup_scatter_kernel(u32 n, const u32* src, const u64* ids, u64* gmap) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) gmap[src[j]] = ids[j];
}
__global__ void dedup_map_roots_kernel(u32 n, const u64* roots, const u64* gmap, u64* out) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = roots[j] ? gmap[id_index(roots[j])] : 0;
}

// ------------------------------------------------------------------------------------------------
// Owner side: intern `n` records into the shard.  Thread-per-key, warp-converged probe loop with
// in-loop publication (see intern_block in vx_build.cuh).  Leaves: record word 8 is the value.
// Shard nodes carry no refcounts (their children live on other GPUs): the merged DAG is read-only.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void intern_records_kernel(InternerDev in, u32 n, const u64* records, u32 rank, bool leaf_round, u64* ids_out,
                                      u32* created) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool need = j < n;
    u64 ch[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u64 val = 0;
    if (need) {
        const u64* r = records + size_t(j) * DEDUP_REC_WORDS;
        for (int k = 0; k < 8; ++k) ch[k] = r[k];
        val = r[8];
    }
    u64 result = 0;
    bool done = !need;
    if (leaf_round) {
        // get_or_create_leaf keyed on the value: u8 -> direct table, wider -> open addressing
        const u32 v = u32(val);
        u32 s = u32(leaf_hash(v)) & in.leaf_mask;
        const u64 mykey = u64(v) | (1ull << 32);
        while (__any_sync(FULL, !done)) {
            if (!done) {
                u64* keyp = sizeof(T) == 1 ? &in.leaf_u8[v] : &in.leaf_keys[s];
                if (sizeof(T) == 1) {
                    u64 g = ld_strong(keyp);
                    if (g == 0) {
                        if (atomicCAS((ull*)keyp, 0ull, (ull)ID_PENDING) == 0ull) {
                            u32 idx = atomicAdd(in.next_index, 1u);
                            if (idx >= in.capacity) {
                                set_error(in, ERR_OOM);
                                st_strong(keyp, 0);
                            } else {
                                ((u8*)in.values)[idx] = u8(v);
                                in.hashes[idx] = leaf_hash(v);
                                for (int q = 0; q < 8; ++q) in.children[size_t(idx) * 8 + q] = 0;  // a leaf has no children row
                                result = id_leaf(u64(idx));
                                fence_release_gpu();
                                st_strong(keyp, result);
                                atomicAdd(created, 1u);
                            }
                            done = true;
                        }
                    } else if (g != ID_PENDING) {
                        result = g;
                        done = true;
                    }
                } else {
                    u64 k = ld_strong(keyp);
                    if (k == 0) {
                        u64 old = atomicCAS((ull*)keyp, 0ull, (ull)mykey);
                        if (old == 0) {
                            u32 idx = atomicAdd(in.next_index, 1u);
                            if (idx >= in.capacity) {
                                set_error(in, ERR_OOM);
                                st_strong(&in.leaf_ids[s], ID_PENDING);  // waiters for this value must not spin for ever
                            } else {
                                ((u32*)in.values)[idx] = v;
                                in.hashes[idx] = leaf_hash(v);
                                for (int q = 0; q < 8; ++q) in.children[size_t(idx) * 8 + q] = 0;  // a leaf has no children row
                                result = id_leaf(u64(idx));
                                fence_release_gpu();
                                st_strong(&in.leaf_ids[s], result);
                                atomicAdd(created, 1u);
                            }
                            done = true;
                        } else if (old != mykey) {
                            s = (s + 1) & in.leaf_mask;
                        }
                    } else if (k == mykey) {
                        u64 g = ld_strong(&in.leaf_ids[s]);
                        if (g != 0) {
                            result = g == ID_PENDING ? 0 : g;  // ID_PENDING: the creator ran out of memory
                            done = true;
                        }
                    } else {
                        s = (s + 1) & in.leaf_mask;
                    }
                }
            }
        }
        if (need) ids_out[j] = with_owner(result, rank);
        return;
    }
    u64 h = 0;
    u32 types = 0, mask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        h += child_hash(ch[k], k);
        types |= u32(id_is_leaf(ch[k])) << k;
        mask |= u32(ch[k] != 0) << k;
    }
    h = finish_hash(h);
    const u32 fp = u32(h >> 47);
    u32 bucket = u32(h) & in.bucket_mask;
    u32 skip = 0;
    int guard = 0;
    while (__any_sync(FULL, !done)) {
        bool claimed = false;
        int ek = 0;
        if (!done) {
            const u64* bp = &in.slots[size_t(bucket) * 8];
            u64 sl[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) ld_strong_v2(bp + 2 * q, &sl[2 * q], &sl[2 * q + 1]);
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
                        cand = sv;
                    }
                }
            }
            if (mb) {
                const u64* rp = &in.children[size_t(u32(cand)) * 8];
                u64 r[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) ld_strong_v2(rp + 2 * q, &r[2 * q], &r[2 * q + 1]);
                bool eq = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) eq = eq && r[k] == ch[k];
                if (eq) {
                    result = id_branch(cand, types, mask);
                    done = true;
                } else {
                    skip |= 1u << (__ffs(mb) - 1);
                }
            } else if (pb) {
                // being published: look again
            } else if (eb) {
                ek = __ffs(eb) - 1;
                claimed = atomicCAS((ull*)&in.slots[size_t(bucket) * 8 + ek], 0ull, (ull)((u64(fp) << 47) | IDX_PENDING)) == 0ull;
            } else {
                bucket = (bucket + 1) & in.bucket_mask;
                skip = 0;
                if (++guard > (1 << 22)) {
                    set_error(in, ERR_TABLE_FULL);
                    done = true;
                }
            }
        }
        const u32 cb = __ballot_sync(FULL, claimed);
        if (cb != 0) {
            u32 base = 0;
            if (lane == 0) base = atomicAdd(in.next_index, u32(__popc(cb)));
            base = __shfl_sync(FULL, base, 0);
            const u32 idx = base + __popc(cb & ((1u << lane) - 1));
            if (claimed) {
                const bool oom = idx >= in.capacity;
                if (oom) {
                    set_error(in, ERR_OOM);
                } else {
                    u64* rp = &in.children[size_t(idx) * 8];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<ulonglong2*>(rp + 2 * q) = make_ulonglong2(ch[2 * q], ch[2 * q + 1]);
                    if (sizeof(T) == 1)
                        ((u8*)in.values)[idx] = u8(val);
                    else
                        ((u32*)in.values)[idx] = u32(val);
                    in.hashes[idx] = h;
                    atomicAdd(created, 1u);
                }
                fence_release_gpu();
                st_strong(&in.slots[size_t(bucket) * 8 + ek], oom ? u64(0) : ((u64(fp) << 47) | u64(idx)));
                result = oom ? 0 : id_branch(u64(idx), types, mask);
                done = true;
            }
        }
    }
    if (need) ids_out[j] = with_owner(result, rank);
}

}  // namespace vx

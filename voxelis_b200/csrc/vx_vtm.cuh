// vx_vtm.cuh — VTM payload (VoxModel::serialize, world/voxmodel.rs:177-294) written straight from the
// device pools.
//
// The reference renumbers every node its two pattern maps hold — leaves first, then branches, each sorted
// by pool index (:199-215) — and writes, per leaf, varint(new id) + value (big endian), per branch,
// varint(new id) + mask + varint(new id of every non-empty child) + LOD value (:231-268).  With the pools in
// HBM that is two prefix sums and one pass:
//   classify   node alive (refcount > 0 <=> still in a pattern map) and leaf / branch (a branch has a child)
//   scan       ranks among leaves / branches in index order  ->  new ids
//   sizes      bytes of every record, indexed by new id;  scan  ->  byte offsets
//   write      one thread per node emits its record
// The chunk table (VTC magic, position, varint(new id of the root), world/voxchunk.rs:382-405) is a few
// bytes per chunk and is appended on the host from the gathered root ids.
#pragma once
#include "vx_device.cuh"

namespace vx {

struct VtmArgs {
    const u64* children;  // [n][8]
    const void* values;   // [n] T
    const u32* refs;      // [n]
    u32 n;                // next_index
    u32 vsize;            // sizeof(T)
    u32 *leaf_flag, *branch_flag;  // [n]   1 where the node is an alive leaf / branch
    u32 *leaf_rank, *branch_rank;  // [n]   exclusive scans of the flags
    u32* newid;                    // [n]   0 for dead nodes and for slot 0 (id_map[0] = 0, :186-187)
    u32* sizes;                    // [n+1] record bytes by (new id - 1), zero beyond the live nodes
    u32* offs;                     // [n+1] exclusive scan of sizes
    u8* out;                       // payload: [4 B][leaf records][4 B][branch records]
};

__device__ __forceinline__ u32 varint_len(u32 v) { return v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5; }
__device__ __forceinline__ u8* put_varint(u8* p, u32 v) {  // io/varint.rs:5-17
    while (v >= 0x80) {
        *p++ = u8((v & 0x7F) | 0x80);
        v >>= 7;
    }
    *p++ = u8(v);
    return p;
}

__global__ void vtm_classify_kernel(VtmArgs a) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    u32 leaf = 0, branch = 0;
    if (i != 0 && a.refs[i] != 0) {
        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(a.children + size_t(i) * 8);
        u64 any = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const ulonglong2 q = row[k];
            any |= q.x | q.y;
        }
        branch = any != 0;
        leaf = !branch;
    }
    a.leaf_flag[i] = leaf;
    a.branch_flag[i] = branch;
}

// new ids (:202-215) and record sizes; `sizes` was zeroed
__global__ void vtm_newid_kernel(VtmArgs a) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const u32 n_leaves = a.leaf_rank[a.n - 1] + a.leaf_flag[a.n - 1];
    u32 id = 0;
    if (a.leaf_flag[i]) id = 1 + a.leaf_rank[i];
    if (a.branch_flag[i]) id = 1 + n_leaves + a.branch_rank[i];
    a.newid[i] = id;
}
__global__ void vtm_sizes_kernel(VtmArgs a) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const u32 id = a.newid[i];
    if (id == 0) return;
    u32 bytes = varint_len(id) + a.vsize;
    if (a.branch_flag[i]) {
        bytes += 1;
        for (int k = 0; k < 8; ++k) {
            const u64 ch = a.children[size_t(i) * 8 + k];
            if (ch != 0) bytes += varint_len(a.newid[id_index(ch)]);
        }
    }
    a.sizes[id - 1] = bytes;
}
__global__ void vtm_write_kernel(VtmArgs a) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const u32 id = a.newid[i];
    if (id == 0) return;
    const bool branch = a.branch_flag[i] != 0;
    u8* p = a.out + (branch ? 8 : 4) + a.offs[id - 1];
    p = put_varint(p, id);
    if (branch) {
        u32 mask = 0;
        u64 ch[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            ch[k] = a.children[size_t(i) * 8 + k];
            mask |= u32(ch[k] != 0) << k;
        }
        *p++ = u8(mask);  // BlockId::mask of a branch == its non-empty children (core/block_id.rs:216-226)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (ch[k] != 0) p = put_varint(p, a.newid[id_index(ch[k])]);
    }
    if (a.vsize == 1) {
        *p = reinterpret_cast<const u8*>(a.values)[i];
    } else {  // to_be_bytes (core/voxel.rs:26-28)
        const u32 v = reinterpret_cast<const u32*>(a.values)[i];
        p[0] = u8(v >> 24), p[1] = u8(v >> 16), p[2] = u8(v >> 8), p[3] = u8(v);
    }
}
__global__ void vtm_roots_kernel(const u32* newid, const u64* roots, u32 n, u32* out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = roots[i] == 0 ? 0u : newid[id_index(roots[i])];
}

// ---- import: VoxModel::deserialize (world/voxmodel.rs:296-408) into a FRESH interner.  The reference gives
// node k of the file pool index k (interner/mod.rs:933,948 assert it), so the host parses the records into
// pool-shaped arrays — children rows as BlockIds, values, in-degree refcounts — copies them in place, and
// this kernel does what deserialize_leaf / deserialize_branch do besides storing: hash and enter every node
// into the tables (mod.rs:941-1000).  One thread per node.
template <class T>
__global__ void vtm_install_kernel(InternerDev in, u32 n_leaves, u32 n_nodes) {
    const u32 idx = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > n_nodes) return;
    if (idx <= n_leaves) {
        const u32 v = sizeof(T) == 1 ? u32(reinterpret_cast<const u8*>(in.values)[idx]) : reinterpret_cast<const u32*>(in.values)[idx];
        in.hashes[idx] = leaf_hash(v);
        const u64 id = id_leaf(idx);
        if (sizeof(T) == 1) {
            in.leaf_u8[v] = id;
        } else {
            const u64 mykey = u64(v) | (1ull << 32);
            u32 s = u32(leaf_hash(v)) & in.leaf_mask;
            for (u32 guard = 0; guard <= in.leaf_mask; ++guard) {
                const u64 old = atomicCAS((ull*)&in.leaf_keys[s], 0ull, (ull)mykey);
                if (old == 0 || old == mykey) {
                    in.leaf_ids[s] = id;
                    return;
                }
                s = (s + 1) & in.leaf_mask;
            }
            set_error(in, ERR_TABLE_FULL);
        }
        return;
    }
    u64 h = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) h += child_hash(in.children[size_t(idx) * 8 + i], i);
    h = finish_hash(h);
    in.hashes[idx] = h;
    const u64 slot = (u64(u32(h >> 47)) << 47) | idx;  // generation 0
    u32 bucket = u32(h) & in.bucket_mask;
    for (u32 guard = 0; guard <= in.bucket_mask; ++guard) {
        for (int k = 0; k < 8; ++k)
            if (atomicCAS((ull*)&in.slots[size_t(bucket) * 8 + k], 0ull, (ull)slot) == 0) return;
        bucket = (bucket + 1) & in.bucket_mask;
    }
    set_error(in, ERR_TABLE_FULL);
}

}  // namespace vx

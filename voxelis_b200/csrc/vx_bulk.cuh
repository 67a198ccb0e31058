// vx_bulk.cuh — level-synchronous bulk builder for MANY fresh trees (sm_100a).
//
// Same result as apply_kernel (vx_build.cuh) for the north-star shape — n fresh trees, no fill flags,
// depth >= 4 — i.e. set_batch_at_depth_iterative (spatial/voxtree.rs:724-1118) on EMPTY roots, which
// reduces to the canonical bottom-up DAG of each batch.  The fused kernel gives every warp
// one 16^3 sub-cube and walks its levels one after the other, so a warp has a handful of dependent
// table lookups in flight; with 10^5..10^8 non-empty blocks in a call that chain latency is the
// bound.  Here every LEVEL of every chunk is one launch with one THREAD per node, so hundreds of
// thousands of independent lookups are in flight and the levels cost a few round trips each:
//
//   plan     warp per 512-block unit reads its 1 KiB of masks once and emits, order-preserving,
//            the candidate blocks (set_mask != 0) and the candidate nodes of the three levels above
//            them as (first child slot, child mask) — siblings are adjacent, so a node's children are
//            a contiguous run of the level below.  Lists are allocated per unit with atomic counters.
//   blocks   thread per candidate block: block_node() of vx_build.cuh (phase 1, :770-897)
//   levels   thread per candidate node: uniform collapse / get_or_create_branch (phase 2, :905-1106)
//   upper    dense levels from the unit nodes (depth D-4) up to the root, thread per node; the last
//            one writes the roots (apply_batch, voxtree.rs:303-328)
//
// HBM layout of the scratch (allocated once per interner for the largest call seen, vx_capi.cu):
//   level l in {0,1,2}:  first[l][cnt_l] u32, cm[l][cnt_l] u8, ids[l][cnt_l] u64   (level 0: first = block index,
//                        cm = set_mask)
//   units:               first[U] u32, cm[U] u8 (dense, U = n * B / 512);  dense id arrays ping-pong above
#pragma once
#include "vx_build.cuh"

namespace vx {

struct BulkArgs {
    InternerDev in;
    const u8* masks;     // [n][B][2]
    const void* values;  // [n][B][8]
    u64* roots;          // [n]
    u8* changed;         // [n] or null
    u32* cnt;            // [0..2] candidates per sparse level, [3] dense units, [4] dense-unit work counter,
                         // [5] non-empty groups of eight units; zeroed before the plan kernel
    u32* first[3];
    u8* cm[3];
    u64* ids[3];
    u32* unit_first;  // [U]
    u8* unit_cm;      // [U]
    u64* dense[2];    // ping-pong dense id arrays for the levels above the units
    u32* dense_units; // [<= U] units with >= dense_min candidate blocks: built by one warp each
    u32* cube_flag;   // [U/8] == epoch: some unit of this group of eight is not empty (no per-call clearing)
    u32 epoch;        // call counter of the interner's bulk scratch, never 0
    u32* cube_list;   // [<= U/8] those groups, in the order they were first seen (count in cnt[5])
    const u32* unit_list;  // optional: the only units that can hold a set bit (host batches know, vx_stage.cuh);
    u32 n_listed;          // the others are neither read nor planned (their unit_cm / roots were zeroed by memsets)
    unsigned long long units;  // U
    u32 n;
    u32 depth;
    u32 blocks;       // B
    u32 dense_min;
    u32 tpk_only;
    u32 use_free;
    // unit memo (busy units only): content digest -> the first unit of THIS call seen with that content
    u64* memo;        // [MEMO_SLOTS][2], own allocation; entries are salted with `epoch`, so old ones never match
    u32* memo_count;  // entries inserted since the table was last cleared (persists across calls)
    u32* unit_delta;  // [U][3] leaf_calls / branch_calls / collapsed of a unit built by bulk_dense_units_kernel
    u32 memo_on;
    u32 weak_first;   // first look at a bucket / stored row through L1 (vx_device.cuh: ld_weak)
    u32 merge_dense;  // busy units are built by the warps of bulk_blocks_kernel once their share of the block list is done
};
#ifndef VX_BULK_MIN_CTAS
#define VX_BULK_MIN_CTAS 3
#endif
constexpr u32 UNIT_PREBUILT = 0xFFFFFFFFu;  // unit_first marker: the unit's node is already in dense[0]
constexpr u32 UNIT_ALIAS = 0xFFFFFFFEu;     // unit_first marker: dense[0] holds the index of the unit with the same content
constexpr u32 MEMO_SLOTS = 1u << 17;        // 16-byte entries (2 MiB)
constexpr u32 MEMO_CLEAR_AT = MEMO_SLOTS / 4;
constexpr int MEMO_PROBES = 8;

__device__ __forceinline__ u32 ld_stream_u32(const void* p) {
    u32 v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_stream_u8(const void* p) {
    u32 v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// 32-byte streaming load (sm_100: LDG.256): one fully coalesced 1 KiB request per warp
__device__ __forceinline__ void ld_stream_v8(const void* p, uint4* a, uint4* b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w)
                 : "l"(p));
}

// With a unit list the plan kernel does not visit every unit, so what it would have written for the ones it
// skips is zeroed here in one launch: the counters, unit_cm, the roots / changed flags (D = 5) or the dense
// level above the units (D >= 6).
__global__ void __launch_bounds__(256) bulk_zero_kernel(BulkArgs a, u32 dense1_bytes) {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nt = size_t(gridDim.x) * blockDim.x;
    if (t < 8) a.cnt[t] = 0;
    for (size_t i = t; i < a.units; i += nt) a.unit_cm[i] = 0;
    for (size_t i = t; i < a.n; i += nt) {
        a.roots[i] = 0;
        if (a.changed) a.changed[i] = 0;
    }
    for (size_t i = t; i * 8 < dense1_bytes; i += nt) a.dense[1][i] = 0;
}

// ------------------------------------------------------------------------------------------------
// plan: one warp per unit of 512 Morton-consecutive blocks (lane = 16 blocks = 32 B of masks).
// ------------------------------------------------------------------------------------------------
// busy units of a warp wait in the lanes' registers (lane k holds the k-th) and enter the queue with ONE atomic per
// 32 of them: a dense world would otherwise put one same-address atomic per unit on a single L2 slice
__device__ __forceinline__ void plan_flush_busy(const BulkArgs& a, u32 lane, u32 my_busy, u32& n_busy) {
    if (n_busy == 0) return;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(&a.cnt[3], n_busy);
    base = __shfl_sync(FULL, base, 0);
    if (lane < n_busy) a.dense_units[base + lane] = my_busy;
    n_busy = 0;
}

// One NON-EMPTY unit w whose 1 KiB of masks sits in the warp's registers (lane = 16 blocks = q0, q1).
__device__ __forceinline__ void plan_unit(const BulkArgs& a, u32 lane, unsigned long long w, const uint4& q0, const uint4& q1,
                                          u32& my_busy, u32& n_busy) {
    // set_mask bytes (even positions; clear_mask is never read by the reference)
    const u64 sa = u64(__byte_perm(q0.x, q0.y, 0x6420)) | (u64(__byte_perm(q0.z, q0.w, 0x6420)) << 32);
    const u64 sb = u64(__byte_perm(q1.x, q1.y, 0x6420)) | (u64(__byte_perm(q1.z, q1.w, 0x6420)) << 32);
    const u32 bits = nzbytes(sa) | (nzbytes(sb) << 8);  // candidate blocks of this lane
    const u32 pa = (bits & 0xFF) != 0, pb = (bits >> 8) != 0;  // this lane's two level-1 parents
    // exclusive prefix sums over the lanes: candidates at level 0 (r0) and level 1 (r1), packed
    const u32 mine = u32(__popc(bits)) | ((pa + pb) << 16);
    u32 incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(FULL, incl, o);
        if (lane >= u32(o)) incl += t;
    }
    const u32 tot = __shfl_sync(FULL, incl, 31);
    const u32 r0 = (incl - mine) & 0xFFFF, r1 = (incl - mine) >> 16;
    const u32 c0 = tot & 0xFFFF, c1 = tot >> 16;
    if (lane == 0 && a.blocks > UNIT_BLOCKS) {  // the group of eight units this one belongs to has work above it
        const u32 cube = u32(w >> 3);
        if (atomicExch(&a.cube_flag[cube], a.epoch) != a.epoch) a.cube_list[atomicAdd(&a.cnt[5], 1u)] = cube;
    }
    if (c0 >= a.dense_min) {
        // a busy unit: thread-per-block lists would cost more than they save; one warp builds it the
        // way apply_kernel does (lane = block, siblings in neighbouring lanes)
        if (lane == n_busy) my_busy = u32(w);
        if (++n_busy == 32) plan_flush_busy(a, lane, my_busy, n_busy);
        if (lane == 0) {
            a.unit_first[w] = UNIT_PREBUILT;
            a.unit_cm[w] = 0xFF;
        }
        return;
    }
    // level 2: node q = lanes 4q..4q+3 (eight level-1 parents)
    u32 cm2 = (pa | (pb << 1)) << (2 * (lane & 3));
    cm2 |= __shfl_xor_sync(FULL, cm2, 1);
    cm2 |= __shfl_xor_sync(FULL, cm2, 2);
    const u32 b2 = __ballot_sync(FULL, (lane & 3) == 0 && cm2 != 0);  // bits at lanes 0,4,..,28
    const u32 c2 = __popc(b2);
    u32 base = 0;
    if (lane < 3) base = atomicAdd(&a.cnt[lane], lane == 0 ? c0 : lane == 1 ? c1 : c2);
    const u32 base0 = __shfl_sync(FULL, base, 0), base1 = __shfl_sync(FULL, base, 1), base2 = __shfl_sync(FULL, base, 2);
    // level 0 entries: global block index + set_mask
    {
        u32 bb = bits, k = base0 + r0;
        const u32 blk0 = u32(w * UNIT_BLOCKS) + lane * 16;
        while (bb) {
            const int j = __ffs(bb) - 1;
            bb &= bb - 1;
            a.first[0][k] = blk0 + j;
            a.cm[0][k] = u8((j < 8 ? sa >> (8 * j) : sb >> (8 * (j - 8))) & 0xFF);
            ++k;
        }
    }
    // level 1 entries: first child slot in level 0 + which of the eight blocks are candidates
    if (pa) {
        a.first[1][base1 + r1] = base0 + r0;
        a.cm[1][base1 + r1] = u8(bits & 0xFF);
    }
    if (pb) {
        a.first[1][base1 + r1 + pa] = base0 + r0 + __popc(bits & 0xFF);
        a.cm[1][base1 + r1 + pa] = u8(bits >> 8);
    }
    // level 2 entries
    if ((lane & 3) == 0 && cm2 != 0) {
        const u32 k2 = base2 + __popc(b2 & ((1u << lane) - 1));
        a.first[2][k2] = base1 + r1;
        a.cm[2][k2] = u8(cm2);
    }
    // the unit node (dense): children = the unit's level-2 candidates
    if (lane == 0) {
        u32 cm3 = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) cm3 |= ((b2 >> (4 * q)) & 1u) << q;
        a.unit_first[w] = base2;
        a.unit_cm[w] = u8(cm3);
    }
}

__device__ __forceinline__ bool plan_any_set(const uint4& q0, const uint4& q1) {
    return ((q0.x | q0.y | q0.z | q0.w | q1.x | q1.y | q1.z | q1.w) & 0x00FF00FFu) != 0;  // set_mask bytes only
}

// plan: one warp per unit of 512 Morton-consecutive blocks (lane = 16 blocks = 32 B of masks: one LDG.256), the masks of
// the warp's next unit in flight while one is planned.  Measured and dropped (profiles/README.md): a cp.async.bulk +
// mbarrier ring of 4 KiB per warp (76 us against 70), a warp per group of eight units with all 8 KiB in flight (87 us; 98
// registers) — the launch is not short of bytes in flight; what it has on top of the streaming read is the per-unit
// list allocation of the non-empty units.
__global__ void __launch_bounds__(256) bulk_plan_kernel(BulkArgs a) {
    const u32 lane = threadIdx.x & 31;
    const unsigned long long warp = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = (size_t(gridDim.x) * blockDim.x) >> 5;
    const bool listed = a.unit_list != nullptr;
    // the previous call's last launch wiped the unit memo if the count stood at MEMO_CLEAR_AT or more
    if (a.memo_on && blockIdx.x == 0 && threadIdx.x == 0 && *a.memo_count >= MEMO_CLEAR_AT) *a.memo_count = 0;
    u32 my_busy = 0, n_busy = 0;
    // unit by unit, the masks of the next unit of this warp in flight while one is planned
    const unsigned long long total = listed ? a.n_listed : a.units;
    auto unit_at = [&](unsigned long long e) -> unsigned long long { return listed ? a.unit_list[e] : e; };
    uint4 na = make_uint4(0, 0, 0, 0), nb = na;
    unsigned long long nw = 0;
    if (warp < total) {
        nw = unit_at(warp);
        ld_stream_v8(a.masks + nw * (UNIT_BLOCKS * 2) + lane * 32, &na, &nb);
    }
    for (unsigned long long e = warp; e < total; e += nwarps) {
        const uint4 q0 = na, q1 = nb;
        const unsigned long long w = nw;
        if (e + nwarps < total) {
            nw = unit_at(e + nwarps);
            ld_stream_v8(a.masks + nw * (UNIT_BLOCKS * 2) + lane * 32, &na, &nb);
        }
        if (!listed && lane == 0 && (w & 7) == 0 && a.blocks > UNIT_BLOCKS) {
            if (a.blocks == 8 * UNIT_BLOCKS) {
                a.roots[w >> 3] = 0;
                if (a.changed) a.changed[w >> 3] = 0;
            } else {
                a.dense[1][w >> 3] = 0;
            }
        }
        if (!__any_sync(FULL, plan_any_set(q0, q1))) {
            if (lane == 0) a.unit_cm[w] = 0;
            continue;
        }
        plan_unit(a, lane, w, q0, q1, my_busy, n_busy);
    }
    plan_flush_busy(a, lane, my_busy, n_busy);
}

// Values of the eight children of a key (leaf value / branch LOD value; 0 for EMPTY), loaded with plain
// cached loads when the children were created by EARLIER launches (bulk_prologue() has dropped whatever
// this SM's L1 held from before).  STRONG = the children may have been created by another SM during THIS
// launch (second level of bulk_upper_kernel): L1 can hold their line from before the value was written,
// so the loads go to L2 (ld.relaxed.gpu) — a plain load there gave a rare wrong LOD value at D = 6,
// tests/test_gpu_bulk.py::test_bulk_lod_values_are_stable.  u8 packs them into one register pair.
template <class T>
struct ChildValues;
template <>
struct ChildValues<u8> {
    u64 w;
    template <bool STRONG>
    __device__ __forceinline__ void load(const InternerDev& in, const u64 (&ch)[8], bool need) {
        u32 b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u8* p = (const u8*)in.values + id_index(ch[i]);
            b[i] = (need && ch[i] != 0) ? (STRONG ? ld_strong_u8(p) : u32(*p)) : 0;
        }
        w = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) w |= u64(b[i]) << (8 * i);
    }
    __device__ __forceinline__ u32 get(int i) const { return u32(w >> (8 * i)) & 0xFF; }
};
template <>
struct ChildValues<int32_t> {
    u32 v[8];
    template <bool STRONG>
    __device__ __forceinline__ void load(const InternerDev& in, const u64 (&ch)[8], bool need) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u32* p = (const u32*)in.values + id_index(ch[i]);
            v[i] = (need && ch[i] != 0) ? (STRONG ? ld_strong(p) : *p) : 0;
        }
    }
    __device__ __forceinline__ u32 get(int i) const { return v[i]; }
};

// ------------------------------------------------------------------------------------------------
// get_or_create_branch (interner/mod.rs:716-829), thread-per-key with the eight child ids in
// registers — the probing protocol of intern_block with the children given directly.
// ------------------------------------------------------------------------------------------------
template <class T, bool STRONG = false>
__device__ inline u64 intern_node(Ctx<T>& c, bool need, const u64 (&ch)[8], u32 types, u32 mask) {
    const InternerDev& in = c.in;
    if (!__any_sync(FULL, need)) return 0;
    u64 h = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) h += child_hash(ch[i], i);
    h = finish_hash(h);
    const u32 fp = u32(h >> 47);
    u32 bucket = u32(h) & in.bucket_mask;
    u64 result = 0;
    bool done = !need;
    // ---- per-warp parent cache (hot keys must not all go to the same L2 line)
    const u32 ue = u32(h >> 32) & (UC - 1);
    if (need) {
        bool hit = true;
#pragma unroll
        for (int i = 0; i < 8; ++i) hit = hit && c.ws->ukey[ue * 8 + i] == ch[i];
        if (hit) {
            result = c.ws->uval[ue];
            done = result != 0;
            if (done) c.t.local++;
        }
    }
    const bool went_global = !done;
    // The child values (needed for the LOD value if the node turns out to be new) are requested together
    // with the first bucket, not after the claim: in a warp step some lane almost always creates a node,
    // so the step would pay that round trip anyway.  Children were published by earlier launches.
    ChildValues<T> cv;
    cv.template load<STRONG>(in, ch, went_global);
    u32 skip = 0;
    int guard = 0;
    bool first = c.weak_first;
    while (__any_sync(FULL, !done)) {
        bool claimed = false;
        int ek = 0;
        if (!done) {
            const u64* bp = &in.slots[size_t(bucket) * 8];
            c.t.probes++;
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
            if (mb) {
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
                for (int i = 0; i < 8; ++i) eq = eq && r[i] == ch[i];
                if (eq) {
                    result = id_branch(cand, types, mask);
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
        // ---- nodes created in this step: one index allocation per warp, payload, fence, publish
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
                        *reinterpret_cast<ulonglong2*>(rp + 2 * j) = make_ulonglong2(ch[2 * j], ch[2 * j + 1]);
                    // LOD value = mode of the child values (core/voxel.rs:96-141); in-degree of every child
                    u32 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = cv.get(i);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (ch[i] != 0) {
                            if (sizeof(T) == 1 && id_is_leaf(ch[i])) {
                                sts_relaxed(&c.cs->leaf[v[i]], ch[i]);  // cta_finish flushes through this table
                                atomicAdd(&c.cs->leafref[v[i]], 1u);
                            } else {
                                ref_add(c, id_index(ch[i]));
                            }
                        }
                    }
                    ((T*)in.values)[idx] = T(lod_value(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]));
                    in.hashes[idx] = h;
                    c.t.branch_miss++;
                }
                fence_release_gpu();
                // out of memory: hand the slot back (the interner is poisoned, results are discarded)
                st_strong(&in.slots[size_t(bucket) * 8 + ek], oom ? u64(0) : ((u64(fp) << 47) | genidx));
                result = oom ? 0 : id_branch(genidx, types, mask);
                done = true;
            }
        }
        first = false;
    }
    // refresh the parent cache: one writer per entry, so an entry is never a mix of two keys
    const bool wr = went_global && result != 0;
    const u32 wb = __ballot_sync(FULL, wr);
    __syncwarp();
    if (wr) {
        const u32 sm = __match_any_sync(wb, ue);
        if ((__ffs(sm) - 1) == c.lane) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c.ws->ukey[ue * 8 + i] = ch[i];
            c.ws->uval[ue] = result;
        }
    }
    __syncwarp();
    return result;
}

// One parent of phase 2 (voxtree.rs:905-1106) on a fresh tree, one thread: absent if no child entered
// `paths`, the shared Leaf if eight identical leaves (:1050, :1062-1075), else the interned branch.
template <class T, bool STRONG = false>
__device__ inline u64 parent_tpk(Ctx<T>& c, bool active, const u64 (&ch)[8]) {
    u32 pres = 0, leafb = 0;
    bool same = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pres |= u32(ch[i] != 0) << i;
        leafb |= u32(id_is_leaf(ch[i])) << i;
        same = same && ch[i] == ch[0];
    }
    const bool any = active && pres != 0;
    const bool collapse = any && same && id_is_leaf(ch[0]);
    if (any) {
        if (collapse)
            c.t.collapsed++;
        else
            c.t.branch_calls++;
    }
    u64 id = intern_node<T, STRONG>(c, any && !collapse, ch, leafb, pres);
    if (collapse) id = ch[0];
    return any ? id : 0;
}

// The level kernels are launched with programmatic stream serialization: their CTAs may start (and run
// this prologue, which touches only shared memory and the argument block) while the previous launch is
// draining; griddepcontrol.wait then blocks until that launch has completed and its writes are visible.
template <class T>
__device__ __forceinline__ void bulk_prologue(Ctx<T>& c, const BulkArgs& a, unsigned char* smem_raw, bool block_level = true) {
    WarpSmem<T>* ws = reinterpret_cast<WarpSmem<T>*>(smem_raw);
    CtaSmem* csp = reinterpret_cast<CtaSmem*>(smem_raw + sizeof(WarpSmem<T>) * WARPS_PER_CTA);
    smem_init<T>(ws, csp, block_level);
    ctx_init<T>(c, a.in, ws, csp, a.use_free != 0);
    c.weak_first = a.weak_first != 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // This CTA may share its SM with the tail of the previous launch, whose reads can have left lines in L1
    // from before other SMs wrote into them (a line of `values` around a node created later, say).  The
    // acquire side of the fence drops those lines (CCTL.IVALL), once per thread, so the plain cached loads
    // below (ChildValues) see what the previous launches wrote.
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// CTAs of a list kernel that have no element to process leave right after the prologue (the grids are
// sized for the worst case the call allows; the real counts only exist on the device).
__device__ __forceinline__ bool bulk_cta_idle(const u32* cnt, u32 per_cta) {
    return size_t(blockIdx.x) * per_cta >= *cnt;
}
// Every warp takes one CONTIGUOUS span of a list (a multiple of 32 elements): neighbours in the list are
// neighbours in space, so the per-warp caches see the local repetition of the world.
__device__ __forceinline__ u32 bulk_span(u32 cnt) {
    const u32 nwarps = gridDim.x * WARPS_PER_CTA;
    return (((cnt + nwarps - 1) / nwarps) + 31u) & ~31u;
}

// ------------------------------------------------------------------------------------------------
// blocks: thread per candidate block.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void dense_units_loop(Ctx<T>& c, const BulkArgs& a);

template <class T>
__global__ void __launch_bounds__(CTA_THREADS, VX_BULK_MIN_CTAS) bulk_blocks_kernel(BulkArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using V = VT<T>;
    Ctx<T> c;
    bulk_prologue<T>(c, a, smem_raw);
    c.tpk_only = a.tpk_only != 0;
    const u32 cnt = a.cnt[0];
    const u32 span = bulk_span(cnt);
    const bool idle = bulk_cta_idle(&a.cnt[0], WARPS_PER_CTA * span);
    if (idle && !(a.merge_dense && a.cnt[3] != 0)) return;
    const u64 first = u64(blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * span;
    const u32 base0 = u32(min(first, u64(cnt)));
    const u32 end = u32(min(first + span, u64(cnt)));
    constexpr u32 stride = 32;
    // software pipeline: (block index, set_mask) two iterations ahead, the block's values one ahead
    u32 blk1 = 0, set1 = 0, blk2 = 0, set2 = 0;
    typename V::Key vals1 = V::zero();
    if (base0 + c.lane < end) {
        blk1 = ld_stream_u32(a.first[0] + base0 + c.lane);
        set1 = ld_stream_u8(a.cm[0] + base0 + c.lane);
        vals1 = V::load(a.values, blk1);
    }
    if (u64(base0) + stride + c.lane < end) {
        blk2 = ld_stream_u32(a.first[0] + base0 + stride + c.lane);
        set2 = ld_stream_u8(a.cm[0] + base0 + stride + c.lane);
    }
    for (u32 base = base0; base < end; base += stride) {
        // poisoned interner: stop (the host reports it).  The word is requested here and tested at the end of the
        // iteration, so its round trip hides behind the work — every eighth iteration only: every warp of the grid
        // polling ONE address each iteration made that load the slowest of the iteration (15.7 % of the stall samples)
        // (and the warps of a CTA take turns, so that the polls of the grid do not all fall into the same iteration)
        const u32 errw = (c.lane == 0 && (((base - base0) / stride + (threadIdx.x >> 5)) & 7) == 7) ? ld_strong(a.in.error) : u32(ERR_NONE);
        const u32 k = base + c.lane;
        const bool active = k < end;
        const u32 set = set1;
        const typename V::Key vals = vals1;
        blk1 = blk2;
        set1 = set2;
        vals1 = V::zero();
        if (u64(k) + stride < end) vals1 = V::load(a.values, blk1);
        if (u64(k) + 2ull * stride < end) {
            blk2 = ld_stream_u32(a.first[0] + k + 2 * stride);
            set2 = ld_stream_u8(a.cm[0] + k + 2 * stride);
        }
        bool present;
        u64 id = block_node<T>(c, active, vals, set, V::zero(), 0, &present);
        if (active) a.ids[0][k] = present ? id : 0;
        if (__any_sync(FULL, errw != ERR_NONE)) break;
    }
    // busy units (queued by the plan kernel) are independent of the lists: the warps take them as they run out of
    // list blocks, so a dense world costs no launch of its own and a mixed one balances itself
    if (a.merge_dense) {
        c.tpk_only = false;
        dense_units_loop<T>(c, a);
    }
    cta_finish<T>(c);
}

// ------------------------------------------------------------------------------------------------
// Unit memo.  A busy unit's node is a pure function of its 1 KiB of set_masks and 8·sizeof(T)·512 bytes of
// values (fresh tree, no fill), and so are the leaf / branch / collapse counts the reference would have
// run up building it.  Worlds repeat themselves at this granularity (every chunk of a checkerboard, every
// slab of rock, the 255 distinct chunks of the "sum per chunk" benchmark), so the warp that is handed a
// busy unit first streams the unit once, forms a 96-bit content digest on the way, and asks a small
// table whether a unit with that content has been seen in THIS call.  If so the unit becomes an alias of
// that unit (the upper kernel copies its node and its counts); if not the warp enters itself and builds
// the unit the usual way.  A hit costs the 5 KiB read and ~300 instructions per lane instead of 16
// iterations of phase 1 + 73 interned parents.
//   entry (16 B, one 128-bit CAS):  x = digest A (salted with the call's epoch, bit 0 forced),
//                                   y = digest B[63:32] << 32 | unit index of the first unit with that content
//                                   (B is salted too and picks the slot: entries of earlier calls lie elsewhere)
// Two units alias only if both 64-bit mum-hash chains agree (A, and the top half of B): 2^-96 per pair of
// distinct units; the reference itself keys its maps on a 64-bit hash with no key comparison
// (interner/hash.rs:15-38).  Entries of earlier calls carry another salt and never match; the table is
// wiped by the call's last launch once MEMO_CLEAR_AT entries have gone in, and inserts stop at half full.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mum64(u64 a, u64 b) { return (a * b) ^ __umul64hi(a, b); }
struct UnitDigest {
    u64 a, b;
    __device__ __forceinline__ void init(u32 lane) {  // the lane (= position in the unit) seeds both chains
        a = 0xA0761D6478BD642Full ^ (u64(lane + 1) * 0x9E3779B97F4A7C15ull);
        b = 0xE7037ED1A0B428DBull ^ (u64(lane + 1) * 0xD6E8FEB86659FD93ull);
    }
    __device__ __forceinline__ void absorb(u64 x, u64 y) {
        a = mum64(x ^ 0x8EBC6AF09C88C6E3ull, y ^ a);
        b = mum64(y ^ 0x589965CC75374CC3ull, x ^ b);
    }
};
__device__ __forceinline__ bool cas128(u64* p, u64 n0, u64 n1, u64* o0, u64* o1) {  // expects (0, 0)
    asm volatile(
        "{\n .reg .b128 c, n, d;\n mov.b128 c, {%2, %2};\n mov.b128 n, {%3, %4};\n"
        " atom.relaxed.gpu.global.cas.b128 d, [%5], c, n;\n mov.b128 {%0, %1}, d;\n}"
        : "=l"(*o0), "=l"(*o1) : "l"(0ull), "l"(n0), "l"(n1), "l"(p) : "memory");
    return *o0 == 0 && *o1 == 0;
}
// lane 0 only.  Returns 1 = alias of unit *rep, 0 = this unit is the first with its content.
__device__ __forceinline__ int memo_lookup(const BulkArgs& a, u64 A, u64 B, u32 w, bool may_insert, u32* rep) {
    const u64 hi = B & 0xFFFFFFFF00000000ull;
    u32 slot = u32(B) & (MEMO_SLOTS - 1);
#pragma unroll 1
    for (int p = 0; p < MEMO_PROBES; ++p) {
        u64* e = a.memo + size_t(slot) * 2;
        u64 x, y;
        ld_strong_v2(e, &x, &y);
        if (x == 0 && y == 0) {
            if (!may_insert) return 0;
            if (cas128(e, A, hi | w, &x, &y)) {
                atomicAdd(a.memo_count, 1u);
                return 0;
            }
        }
        if (x == A && (y & 0xFFFFFFFF00000000ull) == hi) {
            *rep = u32(y);
            return 1;
        }
        slot = (slot + 1) & (MEMO_SLOTS - 1);
    }
    return 0;  // crowded neighbourhood: build it, nobody will alias it
}

// ------------------------------------------------------------------------------------------------
// busy units: one warp per unit.  Stream the unit (masks + values, every byte once), then either: a solid
// unit -> its leaf; content seen before in this call -> alias; else the fused kernel's phase 1 + in-unit
// phase 2 (vx_build.cuh).  The unit's node goes straight into the dense array the upper levels read.
// ------------------------------------------------------------------------------------------------
constexpr u32 DENSE_GRAB = 4;  // units taken from the queue per atomic
template <class T>
__device__ __forceinline__ void dense_units_loop(Ctx<T>& c, const BulkArgs& a) {
    const u32 cnt = a.cnt[3];
    if (cnt == 0) return;
    const u32 upc = a.blocks / UNIT_BLOCKS;
    const int upc_log = 31 - __clz(upc);
    const int D = int(a.depth);
    const u64 salt = mix64(u64(a.epoch) * VX_HC0 + VX_HC1), salt_b = mix64(salt + VX_HC2);
    bool memo_on = a.memo_on != 0;
    bool may_insert = false;
    if (memo_on && c.lane == 0) may_insert = ld_strong(a.memo_count) < MEMO_SLOTS / 2;
    // per-warp memo cache in registers: lane l holds one (A, y) entry.  A world that repeats itself repeats itself
    // nearby, so most units never reach the global table (whose hot entries sit in a single L2 slice).
    u64 cA = 0, cY = 0;
    // units are drawn one at a time while they have to be built (they differ a lot in cost) and DENSE_GRAB at a time
    // while they turn out cheap (solid, or seen before): 10^5 same-address atomics would be the kernel's bound
    bool cheap = false;
    auto grab = [&](u32 k) -> u32 {
        u32 i = 0;
        if (c.lane == 0) i = atomicAdd(&a.cnt[4], k);
        return __shfl_sync(FULL, i, 0);
    };
    u32 take = 1, base = grab(take);
    while (base < cnt) {
        const u32 take_next = cheap ? DENSE_GRAB : 1u;
        const u32 base_next = grab(take_next);  // the next batch is on its way while this one is processed
        if (__any_sync(FULL, c.lane == 0 && ld_strong(a.in.error) != ERR_NONE)) break;
        for (u32 i = base; i < min(base + take, cnt); ++i) {
            cheap = true;
            const u32 w = a.dense_units[i];
            const u32 chunk = w >> upc_log, unit = w & (upc - 1);
            u64 mlo, mhi;
            load_unit_masks(a.masks + size_t(chunk) * a.blocks * 2, size_t(unit) * UNIT_BLOCKS, UNIT_BLOCKS, c.lane, &mlo, &mhi);
            const Under u{0, 0, 0, D};
            const void* cv = (const u8*)a.values + size_t(chunk) * a.blocks * 8 * sizeof(T);
            const uint4* vp = reinterpret_cast<const uint4*>((const u8*)cv + (size_t(unit) * UNIT_BLOCKS + 16 * c.lane) * 8 * sizeof(T));
            constexpr int NV = 8 * int(sizeof(T));  // 16-byte vectors holding this lane's 16 blocks
            const bool full = __all_sync(FULL, (mlo & mhi) == ~0ull);
            // the CTA's warps share what they learn about the world: once misses outnumber hits 4 : 1 nobody asks
            if (memo_on && (i & 3) == 0) {
                const u32 mh = ((volatile u32*)c.cs->memo_stat)[0], mm = ((volatile u32*)c.cs->memo_stat)[1];
                if (mm >= 8 && mh * 4 < mm) memo_on = false;
            }
            if (full || memo_on) {
                // ---- one streaming pass over the unit's values: "solid" test first, then (not solid) the content digest
                uint4 q[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = ld_stream_v4(vp + j);
                const u32 v0 = __shfl_sync(FULL, sizeof(T) == 1 ? (q[0].x & 0xFFu) : q[0].x, 0);
                const u32 splat = sizeof(T) == 1 ? v0 * 0x01010101u : v0;
                bool uni = full && v0 != 0;
                if (uni) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        uni = uni && q[j].x == splat && q[j].y == splat && q[j].z == splat && q[j].w == splat;
#pragma unroll 1
                    for (int j0 = 8; j0 < NV && uni; j0 += 4) {  // wider T: the rest of the lane's 16 blocks
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 r = ld_stream_v4(vp + j0 + j);
                            uni = uni && r.x == splat && r.y == splat && r.z == splat && r.w == splat;
                        }
                    }
                }
                // Solid unit (underground rock, filled volumes): every voxel set to the same non-default value.
                // Phase 1 makes 512 identical leaves, phase 2 collapses 64 + 8 + 1 parents (:826, :1050) — the
                // outcome is that one leaf.
                if (__all_sync(FULL, uni)) {
                    u64 leaf = leaf_get(c, v0, c.lane == 0);
                    if (c.lane == 0) {
                        c.t.leaf_calls += UNIT_BLOCKS;                                    // one per all_same block
                        c.t.collapsed += UNIT_BLOCKS + UNIT_BLOCKS / 8 + UNIT_BLOCKS / 64 + 1;  // blocks + 3 levels
                        a.dense[0][w] = leaf;
                    }
                    continue;
                }
                if (memo_on) {
                    UnitDigest dg;
                    dg.init(c.lane);
                    dg.absorb(mlo, mhi);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dg.absorb(u64(q[j].x) | (u64(q[j].y) << 32), u64(q[j].z) | (u64(q[j].w) << 32));
#pragma unroll 1
                    for (int j0 = 8; j0 < NV; j0 += 4) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 r = ld_stream_v4(vp + j0 + j);
                            dg.absorb(u64(r.x) | (u64(r.y) << 32), u64(r.z) | (u64(r.w) << 32));
                        }
                    }
                    u64 A = mix64(dg.a), B = mix64(dg.b ^ 0x1D8E4E27C47D124Full);
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        A += __shfl_xor_sync(FULL, A, o);
                        B += __shfl_xor_sync(FULL, B, o);
                    }
                    A = (A ^ salt) | 1ull;   // both halves salted with the call: entries of earlier calls lie elsewhere
                    B ^= salt_b;
                    const u64 hi = B & 0xFFFFFFFF00000000ull;
                    int alias = 0;
                    u32 rep = 0;
                    const u32 hit = __ballot_sync(FULL, cA == A && (cY & 0xFFFFFFFF00000000ull) == hi);
                    if (hit) {
                        rep = u32(__shfl_sync(FULL, cY, __ffs(hit) - 1));
                        alias = 1;
                    } else {
                        if (c.lane == 0) alias = memo_lookup(a, A, B, w, may_insert, &rep);
                        alias = __shfl_sync(FULL, alias, 0);
                        rep = alias ? __shfl_sync(FULL, rep, 0) : w;
                        if (c.lane == int((A >> 8) & 31)) {
                            cA = A;
                            cY = hi | rep;
                        }
                    }
                    if (c.lane == 0) atomicAdd(&c.cs->memo_stat[alias ? 0 : 1], 1u);
                    if (alias) {
                        if (c.lane == 0) {
                            a.unit_first[w] = UNIT_ALIAS;
                            a.dense[0][w] = rep;
                        }
                        continue;
                    }
                }
            }
            cheap = false;
            const u32 lc0 = c.t.leaf_calls, bc0 = c.t.branch_calls, co0 = c.t.collapsed;
            const bool some = build_blocks<T, false>(c, mlo, mhi, cv, unit * UNIT_BLOCKS, UNIT_BLOCKS, u);
            u64 node = 0;
            bool present = false;
            if (some) node = reduce_levels<T>(c, UNIT_BLOCKS / 8, D - 2, (unit * UNIT_BLOCKS) >> 3, u, false, &present);
            if (a.memo_on) {  // what an alias of this unit has to add to the counters
                u32 d0 = c.t.leaf_calls - lc0, d1 = c.t.branch_calls - bc0, d2 = c.t.collapsed - co0;
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    d0 += __shfl_xor_sync(FULL, d0, o);
                    d1 += __shfl_xor_sync(FULL, d1, o);
                    d2 += __shfl_xor_sync(FULL, d2, o);
                }
                if (c.lane == 0) {
                    a.unit_delta[size_t(w) * 3 + 0] = d0;
                    a.unit_delta[size_t(w) * 3 + 1] = d1;
                    a.unit_delta[size_t(w) * 3 + 2] = d2;
                }
            }
            if (c.lane == 0) a.dense[0][w] = present ? node : 0;
        }
        base = base_next;
        take = take_next;
    }
}

template <class T>
__global__ void __launch_bounds__(CTA_THREADS, VX_MIN_CTAS) bulk_dense_units_kernel(BulkArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx<T> c;
    bulk_prologue<T>(c, a, smem_raw);
    if (bulk_cta_idle(&a.cnt[3], WARPS_PER_CTA)) return;
    dense_units_loop<T>(c, a);
    cta_finish<T>(c);
}

// ------------------------------------------------------------------------------------------------
// levels 1 and 2 (sparse): thread per candidate node; children = a run of the level below.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(CTA_THREADS, VX_BULK_MIN_CTAS) bulk_level_kernel(BulkArgs a, int level) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx<T> c;
    bulk_prologue<T>(c, a, smem_raw, false);
    const u32 cnt = a.cnt[level];
    const u32 span = bulk_span(cnt);
    if (bulk_cta_idle(&a.cnt[level], WARPS_PER_CTA * span)) return;
    const u32* first = a.first[level];
    const u8* cmv = a.cm[level];
    const u64* below = a.ids[level - 1];
    u64* out = a.ids[level];
    const u64 wfirst = u64(blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * span;
    const u32 base0 = u32(min(wfirst, u64(cnt)));
    const u32 end = u32(min(wfirst + span, u64(cnt)));
    for (u32 base = base0; base < end; base += 32) {
        const u32 errw = (c.lane == 0 && ((((base - base0) >> 5) + (threadIdx.x >> 5)) & 7) == 7) ? ld_strong(a.in.error) : u32(ERR_NONE);  // see bulk_blocks_kernel
        const u32 k = base + c.lane;
        const bool active = k < end;
        const u32 f = active ? ld_stream_u32(first + k) : 0;
        const u32 cm = active ? ld_stream_u8(cmv + k) : 0;
        u64 ch[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            ch[i] = ((cm >> i) & 1) ? ld_stream_u64(below + f + __popc(cm & ((1u << i) - 1))) : 0;
        u64 id = parent_tpk<T>(c, active, ch);
        if (active) out[k] = id;
        if (__any_sync(FULL, errw != ERR_NONE)) break;
    }
    cta_finish<T>(c);
}

// ------------------------------------------------------------------------------------------------
// upper levels (dense): thread per node.  from_units: the nodes are the 512-block units and their
// children come from level 2 through (unit_first, unit_cm); otherwise the children are the eight
// consecutive nodes of the dense array below.  is_root: the nodes are the trees' roots.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void bulk_write_root(Ctx<T>& c, const BulkArgs& a, unsigned long long chunk, u64 id) {
    // apply_batch (voxtree.rs:303-328): nothing entered `paths` -> INVALID -> false, root stays EMPTY
    a.roots[chunk] = id;
    if (a.changed) a.changed[chunk] = id != 0;
    if (id != 0) {  // the tree's root handle; identical chunks share a root: one atomic per distinct root and warp
        const u32 peers = __match_any_sync(__activemask(), id);
        if ((__ffs(peers) - 1) == c.lane) atomicAdd(&c.in.refs[id_index(id)], u32(__popc(peers)));
    }
}

// levels: 1 = the nodes are one level; 2 = lane 0 of every 8-lane group goes on to build the parent of the
// group's eight nodes in the same launch (their ids never leave the registers).  top_is_root: the topmost
// level handled here is the trees' roots; otherwise it is written to `out` (dense).
template <class T>
__global__ void __launch_bounds__(CTA_THREADS, VX_BULK_MIN_CTAS)
bulk_upper_kernel(BulkArgs a, unsigned long long nodes, const u64* below, u64* out, int from_units, int levels,
                  int top_is_root) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx<T> c;
    bulk_prologue<T>(c, a, smem_raw, false);
    const unsigned long long stride = size_t(gridDim.x) * CTA_THREADS;
    const bool poisoned0 = ld_strong(a.in.error) != ERR_NONE;
    // from_units with two levels: only the groups of eight units that hold something are visited (list
    // written by the plan kernel); everything they do not touch was zeroed by the host
    const bool listed = from_units && levels == 2;
    if (listed) nodes = size_t(a.cnt[5]) * 8;
    for (unsigned long long base = (size_t(blockIdx.x) * CTA_THREADS + (threadIdx.x & ~31u)); base < nodes; base += stride) {
        unsigned long long k = base + c.lane;
        const bool active = k < nodes && !poisoned0;
        if (listed) k = k < nodes ? size_t(a.cube_list[k >> 3]) * 8 + (k & 7) : ~0ull;
        u64 ch[8];
        bool prebuilt = false;
        u64 pre_id = 0;
        if (from_units) {
            u32 cm = active ? u32(a.unit_cm[k]) : 0;
            const u32 f = cm ? a.unit_first[k] : 0;
            if (cm && f == UNIT_PREBUILT) {  // built by bulk_dense_units_kernel
                prebuilt = true;
                pre_id = a.dense[0][k];
                cm = 0;
            } else if (cm && f == UNIT_ALIAS) {  // same content as an earlier unit of this call: its node, its counts
                prebuilt = true;
                const u32 rep = u32(a.dense[0][k]);
                pre_id = a.dense[0][rep];
                c.t.leaf_calls += a.unit_delta[size_t(rep) * 3 + 0];
                c.t.branch_calls += a.unit_delta[size_t(rep) * 3 + 1];
                c.t.collapsed += a.unit_delta[size_t(rep) * 3 + 2];
                cm = 0;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                ch[i] = ((cm >> i) & 1) ? ld_stream_u64(a.ids[2] + f + __popc(cm & ((1u << i) - 1))) : 0;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ulonglong2 q = active ? *reinterpret_cast<const ulonglong2*>(below + k * 8 + 2 * j) : make_ulonglong2(0, 0);
                ch[2 * j] = q.x;
                ch[2 * j + 1] = q.y;
            }
        }
        u64 id = parent_tpk<T>(c, active && !prebuilt, ch);
        if (prebuilt) id = pre_id;
        if (levels == 1) {
            if (active || (k < nodes && !listed)) {
                if (top_is_root)
                    bulk_write_root<T>(c, a, k, id);
                else
                    out[k] = id;
            }
            continue;
        }
        // the level above: eight consecutive lanes are siblings (nodes and the warp base are multiples of 8)
#pragma unroll
        for (int i = 0; i < 8; ++i) ch[i] = __shfl_sync(FULL, id, c.gs + i);
        const bool act2 = active && c.li == 0;
        const u64 id2 = parent_tpk<T, true>(c, act2, ch);  // children of this level were created in this launch
        if ((listed ? active : k < nodes) && c.li == 0) {
            if (top_is_root)
                bulk_write_root<T>(c, a, k >> 3, id2);
            else
                out[k >> 3] = id2;
        }
    }
    // the call's last launch wipes the unit memo once enough entries have gone in (the plan kernel of the
    // next call resets the count; nothing touches either in between)
    if (top_is_root && a.memo_on && __shfl_sync(FULL, c.lane == 0 ? ld_strong(a.memo_count) : 0u, 0) >= MEMO_CLEAR_AT) {
        ulonglong2* m = reinterpret_cast<ulonglong2*>(a.memo);
        for (size_t e = size_t(blockIdx.x) * CTA_THREADS + threadIdx.x; e < MEMO_SLOTS; e += stride) m[e] = make_ulonglong2(0, 0);
    }
    cta_finish<T>(c);
}

}  // namespace vx

// vx_capi.cu — host side of the C ABI declared in include/voxelis_b200.h.
//
// Owns all device memory of an interner (SoA pools + tables, see vx_device.cuh), stages host batches
// through double-buffered device slabs, launches the sm_100a kernels of vx_build.cuh / vx_read.cuh and
// turns device-side sticky errors into status codes.  There is no CPU implementation of any compute
// entry point in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/voxelis_b200.h"
#include "vx_build.cuh"
#include "vx_bulk.cuh"
#include "vx_read.cuh"
#include "vx_release.cuh"
#include "vx_dedup.cuh"
#include "vx_stage.cuh"
#include "vx_vtm.cuh"
#include "vx_occupancy.cuh"
#include "vx_terrain.cuh"
#include "vx_voxelize.cuh"

#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <cmath>
#include <unordered_map>

using namespace vx;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(VX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

size_t dtype_size(vx_dtype t) { return t == VX_U8 ? 1 : 4; }
size_t blocks_for_depth(int d) { return size_t(1) << (3 * (d > 0 ? d - 1 : 0)); }

struct Scalars {  // one device allocation for all interner scalars
    u32 next_index;
    u32 free_count;
    u32 error;
    u32 pad;
    Counters ctr;
};

}  // namespace

struct vx_interner {
    int device = 0;
    vx_dtype dtype = VX_U8;
    size_t budget = 0, capacity = 0, nbuckets = 0, leaf_slots = 0;
    InternerDev dev{};
    Scalars* d_scalars = nullptr;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[2]{}, ev_done[2]{};
    int sm_count = 0;
    bool poisoned = false;
    // staging for host-resident batches (double buffered) and small scratch
    void* stage[2]{};
    size_t stage_bytes = 0;
    void* scratch = nullptr;  // device
    size_t scratch_bytes = 0;
    void* hscratch = nullptr;  // pinned host
    size_t hscratch_bytes = 0;
    // single-batch applies (vx_tree_apply_batch): per-call options and results in 64 pinned + mapped bytes that the
    // kernels read / write in place — no small copies either way (layout: struct Mail below)
    void* mail = nullptr;
    uint64_t mail_dev = 0;
    void* mail_stage = nullptr;  // device copy of one small batch (masks + values, <= MAIL_STAGE_BYTES)
    // release (dec_ref_recursive) frontiers, and host mirrors of device state that only changes in
    // synchronous calls
    u64* rel[2]{};
    size_t rel_cap[2]{};
    u32* d_rel_count = nullptr;  // [2]
    void* join = nullptr;        // apply_kernel's join scratch (unit/cube ids + arrival counters)
    size_t join_bytes = 0;
    void* bulk = nullptr;        // bulk builder's level lists (vx_bulk.cuh), sized for the largest call seen
    size_t bulk_bytes = 0;
    uint32_t* bulk_flags = nullptr;  // epoch-tagged "group of units is not empty" flags (own allocation:
    size_t bulk_flags_n = 0;         // they must only ever hold tags, whatever the call sizes were)
    uint32_t bulk_epoch = 0;         // tag of the current call
    uint64_t* bulk_memo = nullptr;   // unit memo of the bulk builder (vx_bulk.cuh): MEMO_SLOTS x 16 B + the insert count
    // diagnostic: CUDA events between the launches of the last apply (vx_interner_profile_stages)
    bool prof = false;
    cudaEvent_t pev[10]{};
    const char* pname[9]{};
    int pstages = 0;
    // diagnostic: phases of the last vx_apply_batches call (host microseconds / device milliseconds)
    std::vector<cudaEvent_t> tevs;
    double trace[8]{};
    bool ever_initialised = false;  // the full clear has run at least once
    bool ever_released = false;     // some node was released since the last full clear (generations / tombstones exist)
    uint64_t free_host = 0;      // entries in the free list
    uint64_t tombs_host = 0;     // deleted table slots since the last rehash
    std::mutex mu;
};

struct vx_tree {
    uint8_t depth;
    bool dirty;
    uint32_t stamp;  // last vx_apply_batches call that listed this tree (duplicate detection)
    vx_block_id root;
};

// Field order: everything vx_apply_batches reads per batch sits in the first cache line (D <= 6).
struct vx_batch {
    uint64_t alias;   // device-visible address of the slot
    int64_t fill;
    uint8_t* masks;   // one slot of the pinned + mapped batch arena: masks[B][2], values[B][8], occ[B/8]
    void* values;
    uint8_t* occ;     // one bit per block: some voxel of the block was set to a non-default value
    // host-side occupancy summary next to has_patches (core/batch.rs:44): one bit per unit of
    // `unit_blocks` Morton-consecutive blocks that Batch::set ever touched.  apply moves only those
    // units across the bus (vx_stage.cuh).
    uint32_t units, unit_blocks;
    uint8_t depth;
    bool has_fill, has_patches;
    bool raw_exposed;     // the caller holds raw array pointers: clear() wipes everything, apply trusts only the masks
    bool journal_ok;      // the journal below is complete (first cache line: vx_apply_batches reads it per batch)
    vx_dtype dtype;
    uint32_t jcount;      // entries in the journal
    uint64_t touched[8];  // <= 512 units (D = 7), inline
    size_t blocks, slot_bytes;
    // Journal (D <= 6, batches only ever written through the API): the blocks that hold a non-default voxel, packed —
    // jblock[k] = block index, jvals[k] = a copy of that block's eight values — in the same pinned slot.  A sparse
    // batch then crosses the bus as 2 + 8·sizeof(T) contiguous bytes per flagged block instead of scattered 8-byte
    // reads that the bus moves as 32-byte sectors (vx_stage.cuh: stage_journal_kernel).  jmap (host only) = slot + 1
    // of a block in the journal.  More than jcap flagged blocks: journal_ok = false and the occupancy path serves.
    uint16_t* jblock;
    void* jvals;
    uint16_t* jmap;
    uint32_t jcap;
};
static_assert(offsetof(vx_batch, touched) <= 64, "per-batch fields of vx_apply_batches' host pass fit one cache line");

namespace {

int ensure_scratch(vx_interner* it, size_t dev_bytes, size_t host_bytes) {
    if (dev_bytes > it->scratch_bytes) {
        if (it->scratch) cudaFree(it->scratch);
        it->scratch = nullptr;
        it->scratch_bytes = 0;
        size_t want = std::max<size_t>(dev_bytes, 1 << 20);
        CU_TRY(cudaMalloc(&it->scratch, want));
        it->scratch_bytes = want;
    }
    if (host_bytes > it->hscratch_bytes) {
        if (it->hscratch) cudaFreeHost(it->hscratch);
        it->hscratch = nullptr;
        it->hscratch_bytes = 0;
        size_t want = std::max<size_t>(host_bytes, 1 << 16);
        CU_TRY(cudaMallocHost(&it->hscratch, want));
        it->hscratch_bytes = want;
    }
    return VX_OK;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Device-visible alias of a pinned + mapped host pointer (cudaMallocHost / cudaHostRegister), else null.
const void* host_device_alias(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    if ((reinterpret_cast<uintptr_t>(a.devicePointer) & 15) != 0) return nullptr;  // the kernel loads 16 B vectors
    return a.devicePointer;
}

int check_device_error(vx_interner* it) {
    u32 err = 0;
    CU_TRY(cudaMemcpyAsync(&err, &it->d_scalars->error, sizeof(u32), cudaMemcpyDeviceToHost, it->stream));
    CU_TRY(cudaStreamSynchronize(it->stream));
    if (err == ERR_NONE) return VX_OK;
    it->poisoned = true;
    if (err == ERR_OOM) return fail(VX_E_OOM, "Out of memory");  // interner/macros.rs:38
    if (err == ERR_TABLE_FULL) return fail(VX_E_OOM, "interner hash table full");
    return fail(VX_E_CUDA, "device-side internal error");
}

// stage markers of the last apply (only when profiling is on)
void prof_begin(vx_interner* it, cudaStream_t s) {
    if (!it->prof) return;
    it->pstages = 0;
    cudaEventRecord(it->pev[0], s);
}
void prof_mark(vx_interner* it, cudaStream_t s, const char* name) {
    if (!it->prof || it->pstages >= 9) return;
    it->pname[it->pstages] = name;
    cudaEventRecord(it->pev[++it->pstages], s);
}

template <class T, bool OLD>
int launch_apply_t(vx_interner* it, int depth, size_t n, const u8* d_masks, const void* d_values, const u8* d_flags,
                   const int64_t* d_fills, const u64* d_old_roots, u64* d_roots, u8* d_changed, cudaStream_t s) {
    if (n == 0) return VX_OK;
    ApplyArgs a{};
    a.in = it->dev;
    a.masks = d_masks;
    a.values = d_values;
    a.flags = d_flags;
    a.fills = (const long long*)d_fills;
    a.old_roots = d_old_roots;
    a.roots = d_roots;
    a.changed = d_changed;
    a.n = u32(n);
    a.depth = u32(depth);
    a.blocks = u32(blocks_for_depth(depth));
    a.use_free = it->free_host > 0 ? 1u : 0u;
    // join scratch (see apply_kernel): arrival counters + dynamic work counter, ids/present flags per
    // unit and per cube
    const size_t upc = a.blocks > UNIT_BLOCKS ? a.blocks / UNIT_BLOCKS : 1, cpc = upc / 8;
    const size_t smem = apply_smem_bytes<T>();
    static thread_local int occ_for_device[64] = {};  // function attributes are per device: set them once each
    int& occ = occ_for_device[it->device & 63];
    if (occ == 0) {
        CU_TRY(cudaFuncSetAttribute(apply_kernel<T, OLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, apply_kernel<T, OLD>, CTA_THREADS, smem));
        occ = std::max(occ, 1);
    }
    const size_t max_ctas = size_t(std::max(occ, 1)) * it->sm_count, max_warps = max_ctas * WARPS_PER_CTA;
    // a warp takes a whole 32^3 cube when there are plenty of cubes; otherwise single units, so that a
    // one-chunk apply still spreads over 8 (D=5) .. 512 (D=7) warps
    a.run = (upc >= 8 && n * cpc >= 4 * max_warps) ? 8u : 1u;
    if (const char* fr = getenv("VX_FORCE_RUN")) {  // tests: exercise both join paths at any size
        if (upc >= 8 && fr[0] == '8') a.run = 8;
        if (fr[0] == '1') a.run = 1;
    }
    const size_t total_runs = (n * upc + a.run - 1) / a.run;
    if (total_runs >= 0xFFFFFFF0ull) return fail(VX_E_INVALID, "too many chunks in one call");
    {
        const size_t counters = (n * cpc + n + 4) * 4;
        const size_t need = counters + n * upc * 8 + n * cpc * 8 + n * upc + n * cpc + 64;
        if (need > it->join_bytes) {
            CU_TRY(cudaStreamSynchronize(s));
            cudaFree(it->join);
            it->join = nullptr;
            it->join_bytes = 0;
            CU_TRY(cudaMalloc(&it->join, need));
            it->join_bytes = need;
        }
        u8* p = (u8*)it->join;
        a.cube_done = (u32*)p;
        a.chunk_done = a.cube_done + n * cpc;
        a.work_next = a.chunk_done + n;
        p += (counters + 15) / 16 * 16;
        a.unit_ids = (u64*)p;
        p += n * upc * 8;
        a.cube_ids = (u64*)p;
        p += n * cpc * 8;
        a.unit_present = p;
        p += n * upc;
        a.cube_present = p;
        CU_TRY(cudaMemsetAsync(it->join, 0, counters, s));
    }
    const size_t ctas = (total_runs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const size_t grid = std::min<size_t>(ctas, max_ctas);
    prof_begin(it, s);
    apply_kernel<T, OLD><<<unsigned(grid), CTA_THREADS, smem, s>>>(a);
    CU_TRY(cudaGetLastError());
    prof_mark(it, s, "apply_kernel");
    if (a.use_free) {
        clamp_free_count_kernel<<<1, 1, 0, s>>>(it->dev);
        CU_TRY(cudaGetLastError());
    }
    return VX_OK;
}

// ---- bulk builder (vx_bulk.cuh): n fresh trees, no flags, depth >= 4 ------------------------------
// candidate blocks (of 512) from which a unit is built by one warp instead of going through the level lists
// 512 = only units whose every block is a candidate: those are the ones that can be solid or repeat an earlier unit
// (the busy-unit loop's two shortcuts).  A partly filled unit is built faster by the level lists — thread per candidate,
// 32 parents per warp step against 4 in the warp-per-unit loop: the surface-and-below terrain 0.903 ms at 128,
// 0.852 at 256, 0.794 at 384, 0.698 at 512; the surface-only headline does not move (profiles/README.md).
u32 bulk_dense_min() {
    if (const char* e = getenv("VX_BULK_DENSE_MIN")) return u32(std::max(1, atoi(e)));
    return 512;
}
// entries of sparse level l: a unit only enters the lists with fewer than dense_min candidate blocks
size_t bulk_level_entries(size_t nb, int l) {
    if (l == 0) return (nb / UNIT_BLOCKS) * std::min<size_t>(UNIT_BLOCKS, bulk_dense_min() - 1) + 32;
    return nb >> (3 * l);
}
size_t bulk_scratch_bytes(size_t n, size_t blocks) {
    const size_t nb = n * blocks;
    size_t need = 256;
    for (int l = 0; l < 3; ++l) need += bulk_level_entries(nb, l) * 13 + 3 * 256;
    const size_t units = nb / UNIT_BLOCKS;
    need += units * 5 + 2 * 256 + units * 8 + units + 2 * 256 + units * 4 + 256 + units + 2 * 256;
    need += units * 12 + 256;  // unit_delta
    return need;
}
size_t stage_max_bytes() {  // device slab of vx_apply_batches (host batches are staged in slices of this size)
    if (const char* e = getenv("VX_STAGE_MAX_BYTES")) {
        unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= (1ull << 16)) return size_t(v);
    }
    return size_t(4) << 30;
}
size_t stage_slices() {
    if (const char* e = getenv("VX_STAGE_SLICES")) {
        long v = strtol(e, nullptr, 10);
        if (v >= 1 && v <= 64) return size_t(v);
    }
    return 0;  // default: geometric cuts, see vx_apply_batches
}
size_t bulk_max_bytes() {
    if (const char* e = getenv("VX_BULK_MAX_BYTES")) return size_t(strtoull(e, nullptr, 10));
    return size_t(8) << 30;
}

// Launch that may overlap its prologue with the tail of the previous kernel in the stream (the kernel
// itself waits with griddepcontrol.wait before it reads anything that kernel wrote).
template <class... P, class... A>
cudaError_t launch_chained(void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool chained,
                           A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = chained ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, P(args)...);
}

template <class T>
int launch_bulk_t(vx_interner* it, int depth, size_t n, const u8* d_masks, const void* d_values, u64* d_roots,
                  u8* d_changed, cudaStream_t s, const u32* d_unit_list = nullptr, u32 n_listed = 0) {
    const size_t blocks = blocks_for_depth(depth), nb = n * blocks, units = nb / UNIT_BLOCKS;
    const size_t need = bulk_scratch_bytes(n, blocks);
    if (need > it->bulk_bytes) {
        CU_TRY(cudaStreamSynchronize(s));
        CU_TRY(cudaStreamSynchronize(it->stream));
        cudaFree(it->bulk);
        it->bulk = nullptr;
        it->bulk_bytes = 0;
        CU_TRY(cudaMalloc(&it->bulk, need));
        it->bulk_bytes = need;
    }
    if (units / 8 + 1 > it->bulk_flags_n) {
        CU_TRY(cudaStreamSynchronize(s));
        CU_TRY(cudaStreamSynchronize(it->stream));
        cudaFree(it->bulk_flags);
        it->bulk_flags = nullptr;
        it->bulk_flags_n = 0;
        CU_TRY(cudaMalloc(&it->bulk_flags, (units / 8 + 1) * 4));
        it->bulk_flags_n = units / 8 + 1;
        CU_TRY(cudaMemsetAsync(it->bulk_flags, 0, it->bulk_flags_n * 4, s));
        if (it->bulk_memo) CU_TRY(cudaMemsetAsync(it->bulk_memo, 0, size_t(MEMO_SLOTS) * 16 + 256, s));  // epochs restart
        it->bulk_epoch = 0;
    }
    const size_t memo_bytes = size_t(MEMO_SLOTS) * 16 + 256;
    if (!it->bulk_memo) {
        CU_TRY(cudaMalloc(&it->bulk_memo, memo_bytes));
        CU_TRY(cudaMemsetAsync(it->bulk_memo, 0, memo_bytes, s));
    }
    if (++it->bulk_epoch == 0) {  // wrapped: the tags of 2^32 calls ago could alias
        CU_TRY(cudaMemsetAsync(it->bulk_flags, 0, it->bulk_flags_n * 4, s));
        CU_TRY(cudaMemsetAsync(it->bulk_memo, 0, memo_bytes, s));
        it->bulk_epoch = 1;
    }
    BulkArgs a{};
    a.in = it->dev;
    a.masks = d_masks;
    a.values = d_values;
    a.roots = d_roots;
    a.changed = d_changed;
    a.units = units;
    a.n = u32(n);
    a.depth = u32(depth);
    a.blocks = u32(blocks);
    a.epoch = it->bulk_epoch;
    a.dense_min = bulk_dense_min();
    a.tpk_only = getenv("VX_BULK_TPK") ? u32(atoi(getenv("VX_BULK_TPK"))) : 0u;
    a.use_free = it->free_host > 0 ? 1u : 0u;
    u8* p = (u8*)it->bulk;
    auto take = [&](size_t bytes) {
        u8* r = p;
        p += (bytes + 255) / 256 * 256;
        return r;
    };
    a.cnt = (u32*)take(32);
    for (int l = 0; l < 3; ++l) {
        const size_t m = bulk_level_entries(nb, l);
        a.ids[l] = (u64*)take(m * 8);
        a.first[l] = (u32*)take(m * 4);
        a.cm[l] = take(m);
    }
    a.unit_first = (u32*)take(units * 4);
    a.unit_cm = take(units);
    a.dense[0] = (u64*)take(units * 8);
    a.dense[1] = (u64*)take(units);
    a.dense_units = (u32*)take(units * 4);
    a.cube_flag = it->bulk_flags;
    a.cube_list = (u32*)take(units / 8 * 4 + 4);
    a.unit_delta = (u32*)take(units * 12);
    a.memo = it->bulk_memo;
    a.memo_count = (u32*)(it->bulk_memo + size_t(MEMO_SLOTS) * 2);
    const char* memo_env = getenv("VX_UNIT_MEMO");  // 0 switches the unit memo off (tests, A/B runs)
    a.memo_on = (memo_env && atoi(memo_env) == 0) ? 0u : 1u;
    const char* weak_env = getenv("VX_WEAK_FIRST");  // 0: every look at the table goes to L2 (A/B runs)
    a.weak_first = (weak_env && atoi(weak_env) == 0) ? 0u : 1u;
    const char* merge_env = getenv("VX_BULK_MERGE_DENSE");  // 0: busy units get a launch of their own (A/B runs)
    a.merge_dense = (merge_env && atoi(merge_env) == 0) ? 0u : 1u;
    a.unit_list = d_unit_list;
    a.n_listed = n_listed;
    prof_begin(it, s);
    if (d_unit_list) {  // what the plan kernel writes for every unit / group it visits, for the ones it will not visit
        const size_t items = std::max(units, n);
        bulk_zero_kernel<<<unsigned(std::min<size_t>((items + 255) / 256, size_t(it->sm_count) * 8)), 256, 0, s>>>(
            a, blocks > 8 * UNIT_BLOCKS ? u32(units) : 0u);
        CU_TRY(cudaGetLastError());
    } else {
        CU_TRY(cudaMemsetAsync(a.cnt, 0, 32, s));
    }
    const size_t smem = apply_smem_bytes<T>();
    static thread_local int occ_for_device[64] = {};  // function attributes are per device: set them once each
    int& occ = occ_for_device[it->device & 63];
    if (occ == 0) {
        CU_TRY(cudaFuncSetAttribute(bulk_blocks_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        CU_TRY(cudaFuncSetAttribute(bulk_level_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        CU_TRY(cudaFuncSetAttribute(bulk_upper_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        CU_TRY(cudaFuncSetAttribute(bulk_dense_units_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bulk_level_kernel<T>, CTA_THREADS, smem));
        occ = std::max(occ, 1);
    }
    const size_t max_ctas = size_t(std::max(occ, 1)) * it->sm_count;
    auto grid_for = [&](size_t threads) {
        return unsigned(std::max<size_t>(1, std::min<size_t>((threads + CTA_THREADS - 1) / CTA_THREADS, max_ctas)));
    };
    {   // plan: a warp per unit, enough warps to keep the memory system full
        // ~8 units per warp: short CTAs that the hardware scheduler balances (units differ a lot in cost)
        static const size_t upw = getenv("VX_PLAN_UPW") ? size_t(atoi(getenv("VX_PLAN_UPW"))) : 8;
        const size_t pu = d_unit_list ? size_t(n_listed) : units;
        const size_t ctas = std::max<size_t>((pu + 8 * upw - 1) / (8 * upw), std::min<size_t>((pu + 7) / 8, size_t(it->sm_count) * 8));
        bulk_plan_kernel<<<unsigned(std::max<size_t>(ctas, 1)), 256, 0, s>>>(a);
        CU_TRY(cudaGetLastError());
        prof_mark(it, s, "bulk_plan_kernel");
    }
    // the sparse levels are sized by device-side counters: launch for the worst case the call allows,
    // capped at one resident wave (the kernels are grid-stride loops)
    static const bool chained = getenv("VX_BULK_NO_PDL") == nullptr;
    CU_TRY(launch_chained(bulk_blocks_kernel<T>, grid_for(nb), CTA_THREADS, smem, s, chained, a));
    prof_mark(it, s, a.merge_dense ? "bulk_blocks_kernel[+busy units]" : "bulk_blocks_kernel");
    CU_TRY(launch_chained(bulk_level_kernel<T>, grid_for(nb >> 3), CTA_THREADS, smem, s, chained, a, 1));
    prof_mark(it, s, "bulk_level_kernel[1]");
    CU_TRY(launch_chained(bulk_level_kernel<T>, grid_for(nb >> 6), CTA_THREADS, smem, s, chained, a, 2));
    prof_mark(it, s, "bulk_level_kernel[2]");
    if (!a.merge_dense) {
        CU_TRY(launch_chained(bulk_dense_units_kernel<T>, grid_for(units * 32), CTA_THREADS, smem, s, chained, a));
        prof_mark(it, s, "bulk_dense_units_kernel");
    }
    // dense levels: units (depth D-4) up to the roots, two levels per launch where there are two
    size_t nodes = units;
    int pp = 1;  // dense[0] holds the prebuilt unit nodes
    const u64* below = nullptr;
    bool from_units = true;
    for (;;) {
        const int levels = nodes == n ? 1 : 2;
        const size_t top = levels == 1 ? nodes : nodes / 8;
        const bool top_is_root = top == n;
        CU_TRY(launch_chained(bulk_upper_kernel<T>, grid_for(nodes), CTA_THREADS, smem, s, chained, a,
                              (unsigned long long)nodes, below, a.dense[pp], from_units ? 1 : 0, levels, top_is_root ? 1 : 0));
        prof_mark(it, s, from_units ? (top_is_root ? "bulk_upper_kernel[units..root]" : "bulk_upper_kernel[units..]")
                                    : (top_is_root ? "bulk_upper_kernel[..root]" : "bulk_upper_kernel"));
        if (top_is_root) break;
        below = a.dense[pp];
        pp ^= 1;
        from_units = false;
        nodes = top / 8;
    }
    if (a.use_free) {
        clamp_free_count_kernel<<<1, 1, 0, s>>>(it->dev);
        CU_TRY(cudaGetLastError());
    }
    return VX_OK;
}

// Which builder takes a call on fresh trees: "bulk" = level-synchronous (vx_bulk.cuh), "fused" =
// apply_kernel.  VX_BUILDER=bulk|fused forces one where it is applicable (tests, A/B runs).
bool use_bulk_builder(int depth, size_t n, const u8* d_masks, const u8* d_flags, const u64* d_old_roots) {
    if (d_flags || d_old_roots || depth < 4) return false;
    if (reinterpret_cast<uintptr_t>(d_masks) & 31) return false;  // the plan kernel reads 32-byte vectors
    const size_t nb = n * blocks_for_depth(depth);
    if (nb >= 0xFFFFFFF0ull) return false;
    if (const char* e = getenv("VX_BUILDER")) {
        if (!strcmp(e, "bulk")) return true;
        if (!strcmp(e, "fused")) return false;
    }
    return nb >= (size_t(1) << 16);  // 16 chunks of 32^3 (measured: the pipeline already wins there, profiles/README.md)
}

// d_old_roots == nullptr: every tree is fresh (the north-star path).
int launch_apply(vx_interner* it, int depth, size_t n, const u8* d_masks, const void* d_values, const u8* d_flags,
                 const int64_t* d_fills, u64* d_roots, u8* d_changed, cudaStream_t s,
                 const u64* d_old_roots = nullptr, const u32* d_unit_list = nullptr, u32 n_listed = 0) {
    if (n > 0xFFFFFFFFull) return fail(VX_E_INVALID, "too many chunks in one call");
    if (d_unit_list) {  // the caller checked with bulk_takes_listed(): one bulk call, only the listed units hold anything
        return it->dtype == VX_U8 ? launch_bulk_t<u8>(it, depth, n, d_masks, d_values, d_roots, d_changed, s, d_unit_list, n_listed)
                                  : launch_bulk_t<int32_t>(it, depth, n, d_masks, d_values, d_roots, d_changed, s, d_unit_list, n_listed);
    }
    if (n > 0 && use_bulk_builder(depth, n, d_masks, d_flags, d_old_roots)) {
        // bound the level lists: very large calls go through in slices (any order gives the same DAG)
        const size_t blocks = blocks_for_depth(depth);
        size_t per = n;
        while (per > 1 && bulk_scratch_bytes(per, blocks) > bulk_max_bytes()) per = (per + 1) / 2;
        for (size_t o = 0; o < n; o += per) {
            const size_t m = std::min(per, n - o);
            const u8* mk = d_masks + o * blocks * 2;
            const void* vl = (const u8*)d_values + o * blocks * 8 * dtype_size(it->dtype);
            int rc = it->dtype == VX_U8 ? launch_bulk_t<u8>(it, depth, m, mk, vl, d_roots + o, d_changed ? d_changed + o : nullptr, s)
                                        : launch_bulk_t<int32_t>(it, depth, m, mk, vl, d_roots + o, d_changed ? d_changed + o : nullptr, s);
            if (rc != VX_OK) return rc;
        }
        return VX_OK;
    }
    if (it->dtype == VX_U8) {
        if (d_old_roots)
            return launch_apply_t<u8, true>(it, depth, n, d_masks, d_values, d_flags, d_fills, d_old_roots, d_roots,
                                            d_changed, s);
        return launch_apply_t<u8, false>(it, depth, n, d_masks, d_values, d_flags, d_fills, nullptr, d_roots,
                                         d_changed, s);
    }
    if (d_old_roots)
        return launch_apply_t<int32_t, true>(it, depth, n, d_masks, d_values, d_flags, d_fills, d_old_roots, d_roots,
                                             d_changed, s);
    return launch_apply_t<int32_t, false>(it, depth, n, d_masks, d_values, d_flags, d_fills, nullptr, d_roots,
                                          d_changed, s);
}

// True when a call of this shape goes to the bulk builder in ONE piece, so it can take a list of units.
bool bulk_takes_listed(int depth, size_t n, const u8* d_masks, const u8* d_flags, const u64* d_old_roots) {
    return n > 0 && n <= 0xFFFFFFFFull && use_bulk_builder(depth, n, d_masks, d_flags, d_old_roots) &&
           bulk_scratch_bytes(n, blocks_for_depth(depth)) <= bulk_max_bytes();
}

int valid_depth(int d) { return d >= 2 && d <= 7; }

__global__ void init_scalars_kernel(u32* words, u32 n) {
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) words[i] = i == 0 ? 1u : 0u;  // next_index = 1 (mod.rs:101)
}

int init_state(vx_interner* it, cudaStream_t s, bool sync) {
    // Nothing has ever been released and no error is known: undo only what the nodes in [1, next_index) did to the
    // tables (reset_used_kernel, vx_release.cuh) instead of clearing tables sized for the whole budget.
    static const bool sparse_ok = !(getenv("VX_RESET_FULL") && atoi(getenv("VX_RESET_FULL")) != 0);
    if (sparse_ok && it->ever_initialised && !it->poisoned && it->free_host == 0 && it->tombs_host == 0 && !it->ever_released) {
        const unsigned grid = unsigned(it->sm_count) * 8;
        if (it->dtype == VX_U8)
            reset_used_kernel<u8><<<grid, 256, 0, s>>>(it->dev, it->nbuckets, it->leaf_slots);
        else
            reset_used_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev, it->nbuckets, it->leaf_slots);
        CU_TRY(cudaGetLastError());
        init_scalars_kernel<<<1, 32, 0, s>>>((u32*)it->d_scalars, u32(sizeof(Scalars) / 4));
        CU_TRY(cudaGetLastError());
        if (sync) CU_TRY(cudaStreamSynchronize(s));
        return VX_OK;
    }
    if (!it->ever_initialised) CU_TRY(cudaMemsetAsync(it->dev.children, 0, it->capacity * 64, s));  // once: no garbage rows
    it->ever_initialised = true;
    it->ever_released = false;
    CU_TRY(cudaMemsetAsync(it->dev.slots, 0, it->nbuckets * 64, s));
    CU_TRY(cudaMemsetAsync(it->dev.refs, 0, it->capacity * 4, s));
    CU_TRY(cudaMemsetAsync(it->dev.gens, 0, it->capacity * 2, s));
    CU_TRY(cudaMemsetAsync(it->dev.children, 0, 64, s));               // slot 0: the empty branch
    CU_TRY(cudaMemsetAsync(it->dev.values, 0, dtype_size(it->dtype), s));
    CU_TRY(cudaMemsetAsync(it->dev.hashes, 0, 8, s));
    if (it->dtype == VX_U8)
        CU_TRY(cudaMemsetAsync(it->dev.leaf_u8, 0, 256 * 8, s));
    else {
        CU_TRY(cudaMemsetAsync(it->dev.leaf_keys, 0, it->leaf_slots * 8, s));
        CU_TRY(cudaMemsetAsync(it->dev.leaf_ids, 0, it->leaf_slots * 8, s));
    }
    // (a kernel, not a memcpy from a stack variable: the async form must not read host memory later)
    init_scalars_kernel<<<1, 32, 0, s>>>((u32*)it->d_scalars, u32(sizeof(Scalars) / 4));
    CU_TRY(cudaGetLastError());
    if (sync) CU_TRY(cudaStreamSynchronize(s));
    it->poisoned = false;
    it->free_host = 0;
    it->tombs_host = 0;
    return VX_OK;
}

int refresh_free_count(vx_interner* it) {
    u32 fc = 0;
    CU_TRY(cudaMemcpyAsync(&fc, &it->d_scalars->free_count, 4, cudaMemcpyDeviceToHost, it->stream));
    CU_TRY(cudaStreamSynchronize(it->stream));
    it->free_host = int32_t(fc) > 0 ? fc : 0;
    return VX_OK;
}

// dec_ref_recursive (interner/mod.rs:419-534) for `n` root handles at once; synchronous.
int release_roots(vx_interner* it, const u64* h_roots, size_t n) {
    if (n == 0) return VX_OK;
    it->ever_released = true;  // generations and tombstones may exist from here on: the next reset clears everything
    cudaStream_t s = it->stream;
    auto ensure = [&](int b, size_t entries) -> int {
        if (entries <= it->rel_cap[b]) return VX_OK;
        cudaFree(it->rel[b]);
        it->rel[b] = nullptr;
        it->rel_cap[b] = 0;
        size_t want = std::max<size_t>(entries, 1 << 16);
        CU_TRY(cudaMalloc(&it->rel[b], want * 8));
        it->rel_cap[b] = want;
        return VX_OK;
    };
    if (!it->d_rel_count) CU_TRY(cudaMalloc(&it->d_rel_count, 8));
    int rc = ensure(0, n);
    if (rc != VX_OK) return rc;
    rc = ensure(1, n);
    if (rc != VX_OK) return rc;
    CU_TRY(cudaMemsetAsync(it->d_rel_count, 0, 8, s));
    // stage the root ids in buffer 1, emit the first frontier into buffer 0
    CU_TRY(cudaMemcpyAsync(it->rel[1], h_roots, n * 8, cudaMemcpyHostToDevice, s));
    release_roots_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(it->dev, it->rel[1], u32(n), it->rel[0],
                                                                    &it->d_rel_count[0]);
    CU_TRY(cudaGetLastError());
    int cur = 0;
    uint64_t freed = 0;
    for (;;) {
        u32 count = 0;
        CU_TRY(cudaMemcpyAsync(&count, &it->d_rel_count[cur], 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        if (count == 0) break;
        freed += count;
        int nxt = cur ^ 1;
        rc = ensure(nxt, size_t(count) * 8);
        if (rc != VX_OK) return rc;
        CU_TRY(cudaMemsetAsync(&it->d_rel_count[nxt], 0, 4, s));
        size_t groups = count;
        unsigned grid = unsigned(std::min<size_t>((groups * 8 + 255) / 256, size_t(it->sm_count) * 8));
        if (it->dtype == VX_U8)
            release_level_kernel<u8><<<grid, 256, 0, s>>>(it->dev, it->rel[cur], count, it->rel[nxt],
                                                          &it->d_rel_count[nxt]);
        else
            release_level_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev, it->rel[cur], count, it->rel[nxt],
                                                               &it->d_rel_count[nxt]);
        CU_TRY(cudaGetLastError());
        cur = nxt;
    }
    it->tombs_host += freed;
    rc = refresh_free_count(it);
    if (rc != VX_OK) return rc;
    // rebuild the tables once deleted slots may take a quarter of one of them (tombs_host counts every freed node,
    // so it bounds the tombstones of either table from above)
    size_t rehash_at = it->nbuckets * 2;
    if (it->dtype != VX_U8) rehash_at = std::min(rehash_at, it->leaf_slots / 4);
    if (const char* e = getenv("VX_REHASH_AT")) rehash_at = size_t(strtoull(e, nullptr, 10));  // tests
    if (it->tombs_host > rehash_at) {
        Scalars sc{};
        CU_TRY(cudaMemcpyAsync(&sc, it->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        u32 next = std::min<u32>(sc.next_index, u32(it->capacity));
        CU_TRY(cudaMemsetAsync(it->dev.slots, 0, it->nbuckets * 64, s));
        if (it->dtype == VX_U8) {
            rehash_kernel<u8><<<(next + 255) / 256, 256, 0, s>>>(it->dev, next);
        } else {
            CU_TRY(cudaMemsetAsync(it->dev.leaf_keys, 0, it->leaf_slots * 8, s));
            CU_TRY(cudaMemsetAsync(it->dev.leaf_ids, 0, it->leaf_slots * 8, s));
            rehash_kernel<int32_t><<<(next + 255) / 256, 256, 0, s>>>(it->dev, next);
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaStreamSynchronize(s));
        it->tombs_host = 0;
    }
    return VX_OK;
}

}  // namespace

extern "C" {

const char* vx_last_error(void) { return g_err.c_str(); }
int vx_abi_version(void) { return VX_ABI_VERSION; }
int vx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return fail(VX_E_CUDA, "no CUDA device");
    }
    return n;
}

// ------------------------------------------------------------------------------- interner
vx_interner* vx_interner_create(size_t budget, vx_dtype dtype, int device) {
    if (dtype != VX_U8 && dtype != VX_I32) {
        fail(VX_E_INVALID, "unsupported voxel type");
        return nullptr;
    }
    size_t node_size = 78 + dtype_size(dtype);  // interner/mod.rs:158-164
    size_t cap = budget / node_size;
    if (cap == 0) {
        fail(VX_E_BUDGET, "Requested budget is too small");  // mod.rs:63-64
        return nullptr;
    }
    if (cap >= 0xFFFFFFF0ull) {
        fail(VX_E_BUDGET, "Requested budget is too large");  // mod.rs:65-68
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        fail(VX_E_CUDA, "no usable CUDA device (this library has no CPU fallback)");
        return nullptr;
    }
    DeviceGuard g(device);
    vx_interner* it = new vx_interner();
    it->device = device;
    it->dtype = dtype;
    it->budget = budget;
    it->capacity = cap;
    size_t nb = 1024;
    while (nb * 4 < cap) nb <<= 1;  // >= 2 slots per node of capacity (load factor <= 0.5)
    it->nbuckets = nb;
    size_t ls = 1024;
    if (dtype != VX_U8)
        while (ls < cap * 2) ls <<= 1;
    it->leaf_slots = ls;
    auto bail = [&](const char* what) -> vx_interner* {
        fail(VX_E_CUDA, std::string("vx_interner_create: ") + what + ": " + cudaGetErrorString(cudaGetLastError()));
        vx_interner_destroy(it);
        return nullptr;
    };
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail("cudaGetDeviceProperties");
    it->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&it->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
    {
        // the copy stream carries the staging kernels of host batches (vx_stage.cuh): one CTA per SM that must
        // not queue behind the builders' full-occupancy grids, so its CTAs are placed first when slots free up
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        if (cudaStreamCreateWithPriority(&it->copy_stream, cudaStreamNonBlocking, hi_prio) != cudaSuccess) return bail("stream");
    }
    for (int i = 0; i < 2; ++i) {
        if (cudaEventCreateWithFlags(&it->ev_copied[i], cudaEventDisableTiming) != cudaSuccess) return bail("event");
        if (cudaEventCreateWithFlags(&it->ev_done[i], cudaEventDisableTiming) != cudaSuccess) return bail("event");
    }
    InternerDev& d = it->dev;
    if (cudaMalloc(&d.children, cap * 64) != cudaSuccess) return bail("children pool");
    if (cudaMalloc(&d.values, cap * dtype_size(dtype)) != cudaSuccess) return bail("values pool");
    if (cudaMalloc(&d.refs, cap * 4) != cudaSuccess) return bail("ref_counts pool");
    if (cudaMalloc(&d.gens, cap * 2) != cudaSuccess) return bail("generations pool");
    if (cudaMalloc(&d.hashes, cap * 8) != cudaSuccess) return bail("hashes pool");
    if (cudaMalloc(&d.free_list, cap * 4) != cudaSuccess) return bail("free list");
    if (cudaMalloc(&d.slots, nb * 64) != cudaSuccess) return bail("branch table");
    if (dtype == VX_U8) {
        if (cudaMalloc(&d.leaf_u8, 256 * 8) != cudaSuccess) return bail("leaf table");
    } else {
        if (cudaMalloc(&d.leaf_keys, ls * 8) != cudaSuccess) return bail("leaf table");
        if (cudaMalloc(&d.leaf_ids, ls * 8) != cudaSuccess) return bail("leaf table");
    }
    if (cudaMalloc(&it->d_scalars, sizeof(Scalars)) != cudaSuccess) return bail("scalars");
    d.bucket_mask = u32(nb - 1);
    d.leaf_mask = u32(ls - 1);
    d.capacity = u32(cap);
    d.next_index = &it->d_scalars->next_index;
    d.free_count = &it->d_scalars->free_count;
    d.error = &it->d_scalars->error;
    d.ctr = &it->d_scalars->ctr;
    if (init_state(it, it->stream, true) != VX_OK) {
        vx_interner_destroy(it);
        return nullptr;
    }
    return it;
}

void vx_interner_destroy(vx_interner* it) {
    if (!it) return;
    DeviceGuard g(it->device);
    if (it->stream) cudaStreamSynchronize(it->stream);
    InternerDev& d = it->dev;
    cudaFree(d.children);
    cudaFree(d.values);
    cudaFree(d.refs);
    cudaFree(d.gens);
    cudaFree(d.hashes);
    cudaFree(d.free_list);
    cudaFree(d.slots);
    cudaFree(d.leaf_u8);
    cudaFree(d.leaf_keys);
    cudaFree(d.leaf_ids);
    cudaFree(it->d_scalars);
    cudaFree(it->bulk);
    cudaFree(it->bulk_flags);
    cudaFree(it->bulk_memo);
    for (auto& e : it->pev)
        if (e) cudaEventDestroy(e);
    for (auto& e : it->tevs)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        cudaFree(it->stage[i]);
        if (it->ev_copied[i]) cudaEventDestroy(it->ev_copied[i]);
        if (it->ev_done[i]) cudaEventDestroy(it->ev_done[i]);
    }
    cudaFree(it->scratch);
    cudaFree(it->join);
    cudaFree(it->rel[0]);
    cudaFree(it->rel[1]);
    cudaFree(it->d_rel_count);
    if (it->hscratch) cudaFreeHost(it->hscratch);
    if (it->mail) cudaFreeHost(it->mail);
    cudaFree(it->mail_stage);
    if (it->stream) cudaStreamDestroy(it->stream);
    if (it->copy_stream) cudaStreamDestroy(it->copy_stream);
    cudaGetLastError();
    delete it;
}

int vx_interner_reset(vx_interner* it) {
    if (!it) return fail(VX_E_INVALID, "null interner");
    DeviceGuard g(it->device);
    return init_state(it, it->stream, true);
}
int vx_interner_reset_async(vx_interner* it, void* stream) {
    if (!it) return fail(VX_E_INVALID, "null interner");
    DeviceGuard g(it->device);
    return init_state(it, stream ? (cudaStream_t)stream : it->stream, false);
}
size_t vx_interner_capacity(const vx_interner* it) { return it ? it->capacity : 0; }
vx_dtype vx_interner_dtype(const vx_interner* it) { return it ? it->dtype : VX_U8; }
int vx_interner_device(const vx_interner* it) { return it ? it->device : -1; }
void* vx_interner_stream(const vx_interner* it) { return it ? (void*)it->stream : nullptr; }

int vx_interner_sync(vx_interner* it) {
    if (!it) return fail(VX_E_INVALID, "null interner");
    DeviceGuard g(it->device);
    CU_TRY(cudaStreamSynchronize(it->stream));
    return check_device_error(it);
}

static int read_scalars(const vx_interner* cit, Scalars* out) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    DeviceGuard g(it->device);
    CU_TRY(cudaMemcpyAsync(out, it->d_scalars, sizeof(Scalars), cudaMemcpyDeviceToHost, it->stream));
    CU_TRY(cudaStreamSynchronize(it->stream));
    return VX_OK;
}

int64_t vx_interner_next_index(const vx_interner* it) {
    if (!it) return fail(VX_E_INVALID, "null interner");
    Scalars s;
    int rc = read_scalars(it, &s);
    if (rc != VX_OK) return rc;
    return std::min<int64_t>(s.next_index, int64_t(it->capacity));
}

static int read_node_field(const vx_interner* cit, vx_block_id id, const void* base, size_t elt, void* out, size_t n) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it) return fail(VX_E_INVALID, "null interner");
    if (id == VX_BLOCK_INVALID || id_index(id) >= it->capacity) return fail(VX_E_INVALID, "invalid block id");
    DeviceGuard g(it->device);
    CU_TRY(cudaMemcpyAsync(out, (const u8*)base + size_t(id_index(id)) * elt, n, cudaMemcpyDeviceToHost, it->stream));
    CU_TRY(cudaStreamSynchronize(it->stream));
    return VX_OK;
}
int vx_interner_get_ref(const vx_interner* it, vx_block_id id, uint32_t* out) {
    if (!out) return fail(VX_E_INVALID, "null out");
    return it ? read_node_field(it, id, it->dev.refs, 4, out, 4) : fail(VX_E_INVALID, "null interner");
}
int vx_interner_get_value(const vx_interner* it, vx_block_id id, int64_t* out) {
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    if (it->dtype == VX_U8) {
        u8 v = 0;
        int rc = read_node_field(it, id, it->dev.values, 1, &v, 1);
        *out = v;
        return rc;
    }
    int32_t v = 0;
    int rc = read_node_field(it, id, it->dev.values, 4, &v, 4);
    *out = v;
    return rc;
}
int vx_interner_get_children(const vx_interner* it, vx_block_id id, vx_block_id out[8]) {
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    if (id_is_leaf(id)) return fail(VX_E_INVALID, "Cannot get children for value node");
    return read_node_field(it, id, it->dev.children, 64, out, 64);
}

int vx_interner_stats(const vx_interner* it, vx_stats* out) {
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    Scalars s;
    int rc = read_scalars(it, &s);
    if (rc != VX_OK) return rc;
    memset(out, 0, sizeof(*out));
    size_t node_size = 78 + dtype_size(it->dtype);
    u64 next = std::min<u64>(s.next_index, it->capacity);
    u64 freec = s.free_count;
    out->requested_budget = it->budget;
    out->actual_budget = it->capacity * node_size;
    out->node_size = node_size;
    out->nodes_capacity = it->capacity;
    out->allocated_nodes = next;  // includes slot 0, like the reference (mod.rs:117)
    out->recycled_nodes = freec;
    out->alive_nodes = next - freec;
    out->patterns = next - freec;
    out->total_deallocations = s.ctr.recycled;
    out->total_allocations = 1 + s.ctr.leaf_misses + s.ctr.branch_misses;
    out->leaf_cache_misses = s.ctr.leaf_misses;
    out->branch_cache_misses = s.ctr.branch_misses;
    out->leaf_cache_hits = s.ctr.leaf_calls - s.ctr.leaf_misses;
    out->branch_cache_hits = s.ctr.branch_calls - s.ctr.branch_misses;
    out->total_cache_hits = out->leaf_cache_hits + out->branch_cache_hits;
    out->total_cache_misses = out->leaf_cache_misses + out->branch_cache_misses;
    out->collapsed_branches = s.ctr.collapsed;
    out->max_alive_nodes = out->alive_nodes;
    out->max_node_id = next ? next - 1 : 0;
    // leaf_nodes / branch_nodes: live counts need the release counters split by kind; for
    // build-only histories they are the miss counters (+ the permanent empty branch, mod.rs:131)
    out->leaf_nodes = s.ctr.leaf_misses;
    out->branch_nodes = 1 + s.ctr.branch_misses;
    // max_*_ref_count / max_generation are running maxima over transient states in the reference;
    // the in-degree refcount model has no such transients: reported as 0.
    return VX_OK;
}

// diagnostic counters not part of InternerStats: probe_steps, cache_hits_local
// Diagnostics (not part of the reference-facing ABI): per-launch device times of the LAST apply call.
int vx_interner_memory(const vx_interner* it, vx_memory* out) {
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    memset(out, 0, sizeof(*out));
    const size_t esz = dtype_size(it->dtype);
    out->pools_bytes = it->capacity * (64 + esz + 4 + 2 + 8 + 4);  // children, values, refs, gens, hashes, free list
    out->table_bytes = it->nbuckets * 64;
    out->leaf_table_bytes = it->dtype == VX_U8 ? 256 * 8 : it->leaf_slots * 16;
    out->bulk_scratch_bytes = it->bulk_bytes + it->bulk_flags_n * 4 + (it->bulk_memo ? size_t(MEMO_SLOTS) * 16 + 256 : 0);
    out->stage_bytes = it->stage_bytes * ((it->stage[0] ? 1 : 0) + (it->stage[1] ? 1 : 0));
    out->other_scratch_bytes = it->scratch_bytes + it->join_bytes + (it->rel_cap[0] + it->rel_cap[1]) * 8 + sizeof(Scalars);
    out->pinned_host_bytes = it->hscratch_bytes;
    out->total_device_bytes = out->pools_bytes + out->table_bytes + out->leaf_table_bytes + out->bulk_scratch_bytes +
                              out->stage_bytes + out->other_scratch_bytes;
    return VX_OK;
}

// diagnostic: state of the bulk builder's unit memo — out[0] entries inserted since the last wipe, out[1] occupied
// slots, out[2..3] the first occupied entry
int vx_interner_debug_memo(const vx_interner* cit, uint64_t out[4]) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!it->bulk_memo) return VX_OK;
    DeviceGuard g(it->device);
    std::vector<u64> h(size_t(MEMO_SLOTS) * 2 + 1);
    CU_TRY(cudaStreamSynchronize(it->stream));
    CU_TRY(cudaMemcpy(h.data(), it->bulk_memo, h.size() * 8, cudaMemcpyDeviceToHost));
    out[0] = u32(h[size_t(MEMO_SLOTS) * 2]);
    for (size_t e = 0; e < MEMO_SLOTS; ++e)
        if (h[2 * e] | h[2 * e + 1]) {
            if (!out[1]) out[2] = h[2 * e], out[3] = h[2 * e + 1];
            ++out[1];
        }
    return VX_OK;
}

int vx_interner_profile_stages(vx_interner* it, int on) {
    if (!it) return fail(VX_E_INVALID, "null interner");
    DeviceGuard g(it->device);
    if (on && !it->pev[0])
        for (auto& e : it->pev) CU_TRY(cudaEventCreate(&e));
    it->prof = on != 0;
    it->pstages = 0;
    return VX_OK;
}
// ms[i] / names[i] for i < return value (<= 9); synchronises the interner's last launches.
int vx_interner_host_trace(const vx_interner* it, double out[8]) {
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    memcpy(out, it->trace, sizeof(it->trace));
    return VX_OK;
}
int vx_interner_stage_ms(vx_interner* it, float ms[9], const char* names[9]) {
    if (!it || !ms) return fail(VX_E_INVALID, "null argument");
    if (!it->prof) return 0;
    DeviceGuard g(it->device);
    for (int i = 0; i < it->pstages; ++i) {
        CU_TRY(cudaEventSynchronize(it->pev[i + 1]));
        CU_TRY(cudaEventElapsedTime(&ms[i], it->pev[i], it->pev[i + 1]));
        if (names) names[i] = it->pname[i];
    }
    return it->pstages;
}

int vx_interner_debug_counters(const vx_interner* it, uint64_t out[8]) {
    if (!it || !out) return fail(VX_E_INVALID, "null argument");
    Scalars s;
    int rc = read_scalars(it, &s);
    if (rc != VX_OK) return rc;
    out[0] = s.ctr.leaf_calls;
    out[1] = s.ctr.branch_calls;
    out[2] = s.ctr.leaf_misses;
    out[3] = s.ctr.branch_misses;
    out[4] = s.ctr.collapsed;
    out[5] = s.ctr.probe_steps;
    out[6] = s.ctr.cache_hits_local;
    out[7] = s.ctr.recycled;
    return VX_OK;
}

int64_t vx_interner_download(const vx_interner* cit, size_t cap, vx_block_id* children, int64_t* values,
                             uint32_t* refs, uint16_t* gens, uint64_t* hashes) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it) return fail(VX_E_INVALID, "null interner");
    int64_t n = vx_interner_next_index(it);
    if (n < 0) return n;
    if (size_t(n) > cap) return fail(VX_E_INVALID, "download: caller arrays too small");
    DeviceGuard g(it->device);
    cudaStream_t s = it->stream;
    if (children) CU_TRY(cudaMemcpyAsync(children, it->dev.children, size_t(n) * 64, cudaMemcpyDeviceToHost, s));
    if (refs) CU_TRY(cudaMemcpyAsync(refs, it->dev.refs, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
    if (gens) CU_TRY(cudaMemcpyAsync(gens, it->dev.gens, size_t(n) * 2, cudaMemcpyDeviceToHost, s));
    if (hashes) CU_TRY(cudaMemcpyAsync(hashes, it->dev.hashes, size_t(n) * 8, cudaMemcpyDeviceToHost, s));
    std::vector<u8> tmp;
    if (values) {
        tmp.resize(size_t(n) * dtype_size(it->dtype));
        CU_TRY(cudaMemcpyAsync(tmp.data(), it->dev.values, tmp.size(), cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (values) {
        if (it->dtype == VX_U8)
            for (int64_t i = 0; i < n; ++i) values[i] = tmp[i];
        else
            for (int64_t i = 0; i < n; ++i) values[i] = ((const int32_t*)tmp.data())[i];
    }
    return n;
}

#include "vx_capi_batch.inl"
// ------------------------------------------------------------------------------- tree
vx_tree* vx_tree_create(uint8_t max_depth) {
    if (!valid_depth(max_depth)) {
        fail(VX_E_INVALID, "Max depth exceeds allowed limit (supported: 2..7)");  // max_depth.rs:77-83
        return nullptr;
    }
    return new vx_tree{max_depth, false, 0u, VX_BLOCK_EMPTY};
}
void vx_tree_destroy(vx_tree* t) { delete t; }
vx_block_id vx_tree_root_id(const vx_tree* t) { return t ? t->root : VX_BLOCK_INVALID; }
uint8_t vx_tree_max_depth(const vx_tree* t) { return t ? t->depth : 0; }
uint32_t vx_tree_voxels_per_axis(const vx_tree* t) { return t ? 1u << t->depth : 0; }
int vx_tree_is_empty(const vx_tree* t) { return t && t->root == 0; }
int vx_tree_is_leaf(const vx_tree* t) { return t && id_is_leaf(t->root); }
int vx_tree_is_dirty(const vx_tree* t) { return t && t->dirty; }
void vx_tree_mark_dirty(vx_tree* t) {
    if (t) t->dirty = true;
}
void vx_tree_clear_dirty(vx_tree* t) {
    if (t) t->dirty = false;
}
int vx_tree_adopt_root(vx_tree* t, vx_block_id root) {
    if (!t || root == VX_BLOCK_INVALID) return fail(VX_E_INVALID, "null tree or invalid root");
    t->root = root;
    t->dirty = true;
    return VX_OK;
}
int vx_trees_forget(vx_tree* const* trees, size_t n) {
    if (n && !trees) return fail(VX_E_INVALID, "null argument");
    for (size_t i = 0; i < n; ++i)
        if (trees[i]) {
            trees[i]->root = VX_BLOCK_EMPTY;
            trees[i]->dirty = false;
        }
    return VX_OK;
}

// ------------------------------------------------------------------------------- apply
int vx_apply_batches_device(vx_interner* it, uint8_t depth, size_t n, const uint8_t* d_masks, const void* d_values,
                            const uint8_t* d_flags, const int64_t* d_fills, vx_block_id* d_roots, uint8_t* d_changed,
                            void* stream) {
    if (!it || !d_masks || !d_values || !d_roots) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(depth)) return fail(VX_E_INVALID, "max_depth must be in [2,7]");
    if (it->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    if ((reinterpret_cast<uintptr_t>(d_masks) | reinterpret_cast<uintptr_t>(d_values)) & 15)
        return fail(VX_E_INVALID, "device masks/values must be 16-byte aligned");
    DeviceGuard g(it->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : it->stream;
    return launch_apply(it, depth, n, d_masks, d_values, d_flags, d_fills, d_roots, d_changed, s);
}

// ------------------------------------------------------------------------------- device batch generation
int vx_terrain_heights_device(vx_interner* it, uint32_t nx, uint32_t nz, uint64_t seed, uint32_t height, int64_t x0,
                              int64_t z0, int32_t* d_heights, void* stream) {
    if (!it || !d_heights) return fail(VX_E_INVALID, "null argument");
    if (!nx || !nz) return VX_OK;
    if (x0 < 0 || z0 < 0 || height > 65536) return fail(VX_E_INVALID, "x0, z0 >= 0 and height <= 65536");
    if (!is_device_ptr(d_heights)) return fail(VX_E_INVALID, "d_heights must be device memory");
    DeviceGuard g(it->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : it->stream;
    const size_t total = size_t(nx) * nz;
    terrain_heights_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(nx, nz, seed, height, x0, z0, d_heights);
    CU_TRY(cudaGetLastError());
    return VX_OK;
}

int vx_terrain_batches_device(vx_interner* it, uint8_t max_depth, const uint32_t grid[3], const int32_t* d_heights,
                              int surface_only, int materials, uint8_t* d_masks, void* d_values, void* stream) {
    if (!it || !grid || !d_heights || !d_masks || !d_values) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(max_depth)) return fail(VX_E_INVALID, "max_depth must be in [2,7]");
    if (materials != 1 && materials != 3) return fail(VX_E_INVALID, "materials must be 1 or 3 (utils/shapes.rs:273-357)");
    if ((reinterpret_cast<uintptr_t>(d_masks) | reinterpret_cast<uintptr_t>(d_values)) & 15)
        return fail(VX_E_INVALID, "device masks/values must be 16-byte aligned");
    if (!is_device_ptr(d_heights) || !is_device_ptr(d_masks) || !is_device_ptr(d_values))
        return fail(VX_E_INVALID, "heights, masks and values must be device memory");
    const size_t n = size_t(grid[0]) * grid[1] * grid[2];
    if (n == 0) return VX_OK;
    DeviceGuard g(it->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : it->stream;
    const size_t total = n * blocks_for_depth(max_depth);  // >= 8 blocks per chunk: divisible by 4
    if (it->dtype == VX_U8)   // 4 Morton-consecutive blocks per thread: one 32-byte value store each
        terrain_batches_kernel<u8, 4><<<unsigned((total / 4 + 255) / 256), 256, 0, s>>>(
            max_depth, grid[0], grid[1], grid[2], d_heights, surface_only, materials, d_masks, (u8*)d_values);
    else                      // one block per thread: 32 bytes of i32 values each
        terrain_batches_kernel<int32_t, 1><<<unsigned((total + 255) / 256), 256, 0, s>>>(
            max_depth, grid[0], grid[1], grid[2], d_heights, surface_only, materials, d_masks, (int32_t*)d_values);
    CU_TRY(cudaGetLastError());
    return VX_OK;
}

int vx_random_batches_device(vx_interner* it, uint8_t max_depth, size_t n, uint64_t seed_base, uint64_t chunk0, uint32_t k,
                             uint32_t cell, uint8_t* d_masks, void* d_values, void* stream) {
    if (!it || !d_masks || !d_values) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(max_depth)) return fail(VX_E_INVALID, "max_depth must be in [2,7]");
    if (k < 1 || (it->dtype == VX_U8 && k > 255) || cell < 1) return fail(VX_E_INVALID, "1 <= k (<= 255 for u8), cell >= 1");
    if ((reinterpret_cast<uintptr_t>(d_masks) | reinterpret_cast<uintptr_t>(d_values)) & 15)
        return fail(VX_E_INVALID, "device masks/values must be 16-byte aligned");
    if (!is_device_ptr(d_masks) || !is_device_ptr(d_values)) return fail(VX_E_INVALID, "masks and values must be device memory");
    if (n == 0) return VX_OK;
    DeviceGuard g(it->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : it->stream;
    const size_t blocks = blocks_for_depth(max_depth), total = n * blocks;
    const u32 blocks_log = u32(3 * (max_depth - 1));
    const unsigned grid = unsigned((total + 255) / 256);
    if (it->dtype == VX_U8)
        random_batches_kernel<u8><<<grid, 256, 0, s>>>(max_depth, total, blocks_log, seed_base, chunk0, k, cell, d_masks, (u8*)d_values);
    else
        random_batches_kernel<int32_t><<<grid, 256, 0, s>>>(max_depth, total, blocks_log, seed_base, chunk0, k, cell, d_masks,
                                                            (int32_t*)d_values);
    CU_TRY(cudaGetLastError());
    return VX_OK;
}

#include "vx_capi_voxelize.inl"
static int apply_slab_impl(vx_interner* it, uint8_t depth, size_t n, const uint8_t* masks, const void* values,
                           const uint8_t* flags, const int64_t* fills, const vx_block_id* h_old_roots,
                           vx_block_id* roots_out, uint8_t* changed_out) {
    if (!it || !masks || !values || !roots_out) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(depth)) return fail(VX_E_INVALID, "max_depth must be in [2,7]");
    if (it->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    if (n == 0) return VX_OK;
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    const size_t B = blocks_for_depth(depth), mbytes = B * 2, vbytes = B * 8 * dtype_size(it->dtype);
    const bool dev_in = is_device_ptr(masks);
    if (dev_in != is_device_ptr(values)) return fail(VX_E_INVALID, "masks and values must live in the same memory space");
    if (dev_in && ((reinterpret_cast<uintptr_t>(masks) | reinterpret_cast<uintptr_t>(values)) & 15))
        return fail(VX_E_INVALID, "device masks/values must be 16-byte aligned");
    const bool dev_roots = is_device_ptr(roots_out), dev_changed = changed_out && is_device_ptr(changed_out);
    cudaStream_t s = it->stream;
    // per-chunk options and outputs live in device scratch: [roots n*8][fills n*8][old n*8][changed n][flags n]
    size_t need = n * 8 + n * 8 + n * 8 + n + n + 64;
    int rc = ensure_scratch(it, need, 0);
    if (rc != VX_OK) return rc;
    u64* d_roots = dev_roots ? roots_out : (u64*)it->scratch;
    int64_t* d_fills = (int64_t*)((u8*)it->scratch + n * 8);
    u64* d_old = (u64*)((u8*)it->scratch + n * 16);
    u8* d_changed = dev_changed ? changed_out : (u8*)it->scratch + n * 24;
    u8* d_flags = (u8*)it->scratch + n * 25;
    const u64* k_old = nullptr;
    if (h_old_roots) {
        CU_TRY(cudaMemcpyAsync(d_old, h_old_roots, n * 8, cudaMemcpyHostToDevice, s));
        k_old = d_old;
    }
    const u8* k_flags = nullptr;
    const int64_t* k_fills = nullptr;
    if (flags) {
        if (is_device_ptr(flags))
            k_flags = flags;
        else {
            // host flags that only say "has patches" are the default: no per-chunk options needed
            bool plain = true;
            for (size_t i = 0; i < n && plain; ++i) plain = flags[i] == VX_FLAG_PATCHES;
            if (!plain) {
                CU_TRY(cudaMemcpyAsync(d_flags, flags, n, cudaMemcpyHostToDevice, s));
                k_flags = d_flags;
            }
        }
    }
    if (fills) {
        if (is_device_ptr(fills))
            k_fills = fills;
        else {
            CU_TRY(cudaMemcpyAsync(d_fills, fills, n * 8, cudaMemcpyHostToDevice, s));
            k_fills = d_fills;
        }
    }
    // Host batches.  Masks are always streamed H2D by the copy engine into one of two staging slabs while
    // the previous slab is being built.  VALUES of pinned (page-locked, mapped) batches are NOT copied:
    // the kernel reads them in place over PCIe ("zero copy") and only for blocks that have a set bit, so
    // a sparse batch moves a fraction of its bytes across the bus and its values never touch HBM.
    //   VX_HOST_MODE=staged    copy masks and values (the only mode for pageable memory)
    //   VX_HOST_MODE=zerocopy  the kernel reads masks and values in place
    const void* zm = dev_in ? nullptr : host_device_alias(masks);
    const void* zv = dev_in ? nullptr : host_device_alias(values);
    const char* mode = getenv("VX_HOST_MODE");
    const bool pinned = zm && zv;
    const bool all_zero_copy = pinned && mode && strcmp(mode, "zerocopy") == 0;
    const bool copy_values = !pinned || (mode && strcmp(mode, "staged") == 0);
    if (dev_in) {
        rc = launch_apply(it, depth, n, masks, values, k_flags, k_fills, d_roots, d_changed, s, k_old);
        if (rc != VX_OK) return rc;
    } else if (all_zero_copy) {
        rc = launch_apply(it, depth, n, (const u8*)zm, zv, k_flags, k_fills, d_roots, d_changed, s, k_old);
        if (rc != VX_OK) return rc;
    } else {
        size_t per = mbytes + (copy_values ? vbytes : 0);
        size_t slab = std::max<size_t>(1, std::min<size_t>(n, (size_t(copy_values ? 96 : 32) << 20) / per));
        if (slab * per > it->stage_bytes) {
            for (int i = 0; i < 2; ++i) {
                cudaFree(it->stage[i]);
                it->stage[i] = nullptr;
            }
            it->stage_bytes = 0;
            for (int i = 0; i < 2; ++i) CU_TRY(cudaMalloc(&it->stage[i], slab * per));
            it->stage_bytes = slab * per;
        }
        CU_TRY(cudaEventRecord(it->ev_done[0], s));  // orders the flags/fills copies, frees both slabs
        CU_TRY(cudaEventRecord(it->ev_done[1], s));
        size_t k = 0;
        for (size_t lo = 0; lo < n; lo += slab, ++k) {
            size_t cnt = std::min(slab, n - lo);
            int b = int(k & 1);
            u8* sm = (u8*)it->stage[b];
            const void* sv = (const u8*)zv + lo * vbytes;
            CU_TRY(cudaStreamWaitEvent(it->copy_stream, it->ev_done[b], 0));
            CU_TRY(cudaMemcpyAsync(sm, masks + lo * mbytes, cnt * mbytes, cudaMemcpyHostToDevice, it->copy_stream));
            if (copy_values) {
                u8* dv = sm + slab * mbytes;
                CU_TRY(cudaMemcpyAsync(dv, (const u8*)values + lo * vbytes, cnt * vbytes, cudaMemcpyHostToDevice,
                                       it->copy_stream));
                sv = dv;
            }
            CU_TRY(cudaEventRecord(it->ev_copied[b], it->copy_stream));
            CU_TRY(cudaStreamWaitEvent(s, it->ev_copied[b], 0));
            rc = launch_apply(it, depth, cnt, sm, sv, k_flags ? k_flags + lo : nullptr, k_fills ? k_fills + lo : nullptr,
                              d_roots + lo, d_changed + lo, s, k_old ? k_old + lo : nullptr);
            if (rc != VX_OK) return rc;
            CU_TRY(cudaEventRecord(it->ev_done[b], s));
        }
    }
    if (!dev_roots) CU_TRY(cudaMemcpyAsync(roots_out, d_roots, n * 8, cudaMemcpyDeviceToHost, s));
    if (changed_out && !dev_changed) CU_TRY(cudaMemcpyAsync(changed_out, d_changed, n, cudaMemcpyDeviceToHost, s));
    rc = check_device_error(it);
    if (rc == VX_OK && it->free_host > 0) rc = refresh_free_count(it);
    return rc;
}

// One Batch handle on one tree, the latency path (BASELINE config 1: the reference's own apply_batch takes 23 us).
// The batch sits in the pinned + mapped arena.  A small one (D = 5; u8 at D = 6) crosses the bus as ONE copy of its
// contiguous masks + values on the interner's stream — a kernel reading them in place waits out a PCIe round trip per
// 32-block iteration, 44 us for a 32^3 chunk against ~25 with the copy; a big one is read in place (only blocks with a set
// bit move).  The flag / fill / old root come from the interner's mailbox and the root and the changed flag go back
// into it (mapped memory, no copies); the host's one wait is for the error word.
// (apply_slab_impl for the same call: two pageable H2D copies, the masks through the copy stream behind two events, three
// D2H copies each with its own synchronise, six cudaPointerGetAttributes: 80 us.)
constexpr size_t MAIL_STAGE_BYTES = 512 << 10;
struct Mail {
    uint64_t root, old_root;
    int64_t fill;
    uint32_t err;
    uint8_t changed, flag;
};
int apply_one_in_place(vx_interner* it, const vx_batch* b, uint8_t flag, int64_t fill, vx_block_id old_root,
                       vx_block_id* root_out, uint8_t* changed_out) {
    if (it->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    if (!it->mail) {
        void* dev = nullptr;
        CU_TRY(cudaHostAlloc(&it->mail, 64, cudaHostAllocMapped));
        CU_TRY(cudaHostGetDevicePointer(&dev, it->mail, 0));
        it->mail_dev = reinterpret_cast<uint64_t>(dev);
    }
    static_assert(sizeof(Mail) <= 64, "mailbox layout");
    Mail* m = static_cast<Mail*>(it->mail);
    *m = Mail{0, old_root, fill, 0, 0, flag};
    auto dev = [&](size_t off) { return reinterpret_cast<void*>(it->mail_dev + off); };
    const u8* zm = reinterpret_cast<const u8*>(b->alias);
    const size_t voff = size_t(static_cast<const u8*>(b->values) - b->masks);
    const void* zv = zm + voff;
    cudaStream_t s = it->stream;
    const size_t span = voff + b->blocks * 8 * dtype_size(b->dtype);  // masks .. end of values, contiguous in the slot
    if ((flag & VX_FLAG_PATCHES) && span <= MAIL_STAGE_BYTES) {
        if (!it->mail_stage) CU_TRY(cudaMalloc(&it->mail_stage, MAIL_STAGE_BYTES));
        CU_TRY(cudaMemcpyAsync(it->mail_stage, b->masks, span, cudaMemcpyHostToDevice, s));
        zm = static_cast<const u8*>(it->mail_stage);
        zv = zm + voff;
    }
    int rc = launch_apply(it, b->depth, 1, zm, zv, flag == VX_FLAG_PATCHES ? nullptr : static_cast<const u8*>(dev(offsetof(Mail, flag))),
                          (flag & VX_FLAG_FILL) ? static_cast<const int64_t*>(dev(offsetof(Mail, fill))) : nullptr,
                          static_cast<u64*>(dev(offsetof(Mail, root))), static_cast<u8*>(dev(offsetof(Mail, changed))), s,
                          old_root != VX_BLOCK_EMPTY ? static_cast<const u64*>(dev(offsetof(Mail, old_root))) : nullptr);
    if (rc != VX_OK) return rc;
    CU_TRY(cudaMemcpyAsync(&m->err, &it->d_scalars->error, sizeof(u32), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (m->err != ERR_NONE) {
        it->poisoned = true;
        if (m->err == ERR_OOM) return fail(VX_E_OOM, "Out of memory");  // interner/macros.rs:38
        if (m->err == ERR_TABLE_FULL) return fail(VX_E_OOM, "interner hash table full");
        return fail(VX_E_CUDA, "device-side internal error");
    }
    *root_out = m->root;
    *changed_out = m->changed;
    if (it->free_host > 0) return refresh_free_count(it);
    return VX_OK;
}

int vx_apply_batches_slab(vx_interner* it, uint8_t depth, size_t n, const uint8_t* masks, const void* values,
                          const uint8_t* flags, const int64_t* fills, vx_block_id* roots_out, uint8_t* changed_out) {
    return apply_slab_impl(it, depth, n, masks, values, flags, fills, nullptr, roots_out, changed_out);
}

int vx_tree_apply_batch(vx_interner* it, vx_tree* t, const vx_batch* b) {
    if (!it || !t || !b) return fail(VX_E_INVALID, "null argument");
    if (b->depth != t->depth) return fail(VX_E_INVALID, "batch and tree depths differ");
    if (b->dtype != it->dtype) return fail(VX_E_INVALID, "batch and interner voxel types differ");
    uint8_t flag = (b->has_fill ? VX_FLAG_FILL : 0) | (b->has_patches ? VX_FLAG_PATCHES : 0);
    int64_t fill = b->fill;
    vx_block_id root = 0, old_root = t->root;
    uint8_t changed = 0;
    // a non-empty tree is merged with the batch on the device (old-tree descent, voxtree.rs:785-842,
    // :930-952); with a fill the batch is built against Leaf(fill) and the old tree only released
    int rc;
    static const bool lean = getenv("VX_SINGLE_STAGED") == nullptr;
    if (lean && b->alias && !getenv("VX_HOST_MODE"))
        rc = apply_one_in_place(it, b, flag, fill, old_root, &root, &changed);
    else
        rc = apply_slab_impl(it, b->depth, 1, b->masks, b->values, &flag, &fill,
                             old_root != VX_BLOCK_EMPTY ? &old_root : nullptr, &root, &changed);
    if (rc != VX_OK) return rc;
    if (!changed) return 0;
    if (old_root != VX_BLOCK_EMPTY) {  // voxtree.rs:309-314 (the reference asserts new != old here)
        std::lock_guard<std::mutex> lk(it->mu);
        DeviceGuard g(it->device);
        rc = release_roots(it, &old_root, 1);
        if (rc != VX_OK) return rc;
    }
    t->root = root;
    t->dirty = true;
    return 1;
}

int vx_apply_batches(vx_interner* it, vx_tree* const* trees, const vx_batch* const* batches, size_t n,
                     uint8_t* changed) {
    if (!it || (n && (!trees || !batches))) return fail(VX_E_INVALID, "null argument");
    if (n == 0) return VX_OK;
    bool fused = true;
    static std::atomic<uint32_t> g_call{0};
    const uint32_t stamp = ++g_call;  // a tree may appear only once in the fused path: stamp and compare
    if (!trees[0] || !batches[0]) return fail(VX_E_INVALID, "null tree or batch");
    const uint8_t depth0 = batches[0]->depth;
    auto serial = [&]() -> int {  // mixed depths or repeated trees: the reference's serial loop (lib.rs:357-361)
        for (size_t i = 0; i < n; ++i) {
            int rc = vx_tree_apply_batch(it, trees[i], batches[i]);
            if (rc < 0) return rc;
            if (changed) changed[i] = uint8_t(rc);
        }
        return VX_OK;
    };
    if (it->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    std::unique_lock<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    const int depth = depth0;
    const size_t B = blocks_for_depth(depth), mbytes = B * 2, vbytes = B * 8 * dtype_size(it->dtype);
    const size_t per = mbytes + vbytes;
    const uint32_t upc = batches[0]->units, ub = batches[0]->unit_blocks;
    uint32_t upc_log2 = 0;
    while ((1u << upc_log2) < upc) ++upc_log2;
    // The batches sit in pinned host memory, each with its list of touched units.  Only those units cross
    // the bus: per slice, the device slab's masks are zeroed in HBM, stage_units_kernel pulls the touched
    // units' masks and the values of blocks with a set bit (vx_stage.cuh), and the builders run on the slab.
    auto now_us = [] {
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    const double t_begin = now_us();
    // Slices: a big call is cut so that the bus traffic of slice k+1 (copy_stream) overlaps the build of
    // slice k (the interner's stream) and the host's list building for slice k+2; two slabs.  By default
    // the cuts are at n/8 and n/2: a small first slice puts the bus to work early, while the host is still
    // listing the rest (measured on the perlin world: 1.22 ms against 1.30 with a fourth slice at 7n/8, 1.33
    // with two equal slices, 1.44 with one).  VX_STAGE_SLICES=k asks for k equal slices instead.  A third slab, so that
    // the last slice's traffic need not wait for the first build to hand its slab back: 0.965 ms against 0.967 — by the
    // time the host has checked the last handle (~0.43 ms in) the bus has long caught up, and what follows is the three
    // builds back to back (0.50 ms summed: small calls), not the staging.
    size_t slice = std::max<size_t>(1, std::min<size_t>(n, stage_max_bytes() / (2 * per)));
    slice = std::min<size_t>(slice, size_t(0xFFFFFFF0u) >> upc_log2);
    std::vector<size_t> cut{0};
    if (n >= 8192 && stage_slices() == 0 && (n + 1) / 2 <= slice) {
        std::vector<double> fr{0.125, 0.5};
        if (const char* e = getenv("VX_STAGE_CUTS")) {  // experiments: "0.1,0.4,0.8"
            fr.clear();
            for (const char* q = e; *q;) {
                char* end = nullptr;
                double v = strtod(q, &end);
                if (end == q) break;
                if (v > 0 && v < 1) fr.push_back(v);
                q = *end ? end + 1 : end;
            }
        }
        size_t widest = 0;
        for (double f : fr) {
            size_t c = std::min(n - 1, std::max(cut.back() + 1, size_t(f * double(n))));
            widest = std::max(widest, c - cut.back());
            cut.push_back(c);
        }
        widest = std::max(widest, n - cut.back());
        if (widest <= slice) {
            slice = widest;
        } else {  // the cuts asked for do not fit the slabs: equal slices
            cut.assign(1, 0);
            for (size_t lo = slice; lo < n; lo += slice) cut.push_back(lo);
        }
    } else {
        if (n >= 8192 && stage_slices() > 0) slice = std::min(slice, (n + stage_slices() - 1) / stage_slices());
        for (size_t lo = slice; lo < n; lo += slice) cut.push_back(lo);
    }
    cut.push_back(n);
    const size_t n_slices = cut.size() - 1;
    const int n_slabs = n_slices > 1 ? 2 : 1;
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t max_units = n * upc;
    const size_t slab_bytes = up(slice * mbytes) + up(slice * vbytes);
    const size_t o_roots = n_slabs * slab_bytes, o_fills = o_roots + up(n * 8), o_old = o_fills + up(n * 8),
                 o_src = o_old + up(n * 8), o_units = o_src + up(n * 8), o_changed = o_units + up(max_units * 4),
                 o_flags = o_changed + up(n), o_jcnt = o_flags + up(n), dev_need = o_jcnt + up(n * 4);
    const size_t h_fills_o = up(n * 8), h_old_o = h_fills_o + up(n * 8), h_src_o = h_old_o + up(n * 8),
                 h_units_o = h_src_o + up(n * 8), h_changed_o = h_units_o + up(max_units * 4),
                 h_flags_o = h_changed_o + up(n), h_jcnt_o = h_flags_o + up(n), host_need = h_jcnt_o + up(n * 4);
    int rc = ensure_scratch(it, dev_need, host_need);
    if (rc != VX_OK) return rc;
    u8* dbase = (u8*)it->scratch;
    u64* d_roots = (u64*)(dbase + o_roots);
    int64_t* d_fills = (int64_t*)(dbase + o_fills);
    u64* d_old = (u64*)(dbase + o_old);
    u64* d_src = (u64*)(dbase + o_src);
    u32* d_units = (u32*)(dbase + o_units);
    u8* d_changed = dbase + o_changed;
    u8* d_flags = dbase + o_flags;
    u32* d_jcnt = (u32*)(dbase + o_jcnt);
    u8* hbase = (u8*)it->hscratch;
    u64* h_roots = (u64*)hbase;
    int64_t* h_fills = (int64_t*)(hbase + h_fills_o);
    u64* h_old = (u64*)(hbase + h_old_o);
    u64* h_src = (u64*)(hbase + h_src_o);
    u32* h_units = (u32*)(hbase + h_units_o);
    u8* h_changed = hbase + h_changed_o;
    u8* h_flags = hbase + h_flags_o;
    u32* h_jcnt = (u32*)(hbase + h_jcnt_o);
    // where a batch's journal sits in its slot (same for every batch of the call: one depth, one dtype)
    const size_t ob = ((B + 7) / 8 + 3 + 15) & ~size_t(15);
    const size_t off_jvals = mbytes + vbytes + ob, off_jblock = off_jvals + size_t(batches[0]->jcap) * 8 * dtype_size(it->dtype);
    cudaStream_t s = it->stream, cs = it->copy_stream;
    const bool trace = it->prof;
    if (trace) {
        it->tevs.resize(4 * n_slices, nullptr);
        for (auto& e : it->tevs)
            if (!e) CU_TRY(cudaEventCreate(&e));
    }
    // Per slice, on the host: check the handles, fill the per-chunk descriptors and list the touched units, in
    // one pass.  The bus is at work after ~1/8 of that pass; nothing touches the interner before every slice
    // has been checked.
    struct Prep {
        int rc = VX_OK;
        const char* err = nullptr;
        bool fused = true, any_fill = false, old_here = false, raw = false, journal = true;
        size_t nu = 0;
    };
    std::vector<Prep> prep(n_slices);
    auto prepare = [&](size_t sl) {
        Prep& P = prep[sl];
        const size_t lo = cut[sl], hi = cut[sl + 1];
        u32* out = h_units + lo * upc;
        size_t k = 0;
        for (size_t i = lo; i < hi; ++i) {
            if (i + 8 < hi) {  // the handles are separate heap objects: keep the misses in flight
                __builtin_prefetch(trees[i + 8]);
                __builtin_prefetch(batches[i + 8]);
                __builtin_prefetch((const char*)batches[i + 8] + 64);
            }
            vx_tree* t = trees[i];
            const vx_batch* b = batches[i];
            if (!t || !b) {
                P.rc = VX_E_INVALID, P.err = "null tree or batch";
                return;
            }
            if (b->depth != t->depth) {
                P.rc = VX_E_INVALID, P.err = "batch and tree depths differ";
                return;
            }
            if (b->dtype != it->dtype) {
                P.rc = VX_E_INVALID, P.err = "batch and interner voxel types differ";
                return;
            }
            // a tree may appear only once in the fused path: stamp and compare
            if (b->depth != depth0 || t->stamp == stamp) P.fused = false;
            t->stamp = stamp;
            h_flags[i] = (b->has_fill ? VX_FLAG_FILL : 0) | (b->has_patches ? VX_FLAG_PATCHES : 0);
            h_fills[i] = b->fill;
            h_old[i] = t->root;
            h_src[i] = b->alias;
            P.any_fill = P.any_fill || b->has_fill;
            P.old_here = P.old_here || t->root != VX_BLOCK_EMPTY;
            P.raw = P.raw || b->raw_exposed;
            P.journal = P.journal && b->journal_ok;
            h_jcnt[i] = b->has_patches ? b->jcount : 0;
            if (!b->has_patches || !P.fused) continue;
            const u32 first = u32(i - lo) << upc_log2;
            for (uint32_t w = 0; w * 64 < b->units; ++w)
                for (uint64_t bits = b->touched[w]; bits; bits &= bits - 1) out[k++] = first + w * 64 + u32(__builtin_ctzll(bits));
        }
        P.nu = k;
    };
    // (One short-lived thread per later slice was tried for this pass: creating and joining them cost more than
    // the ~0.2 ms they took off the critical path — 1.25-1.5 ms per perlin world against 1.05 ms inline.  Two helper
    // threads that live as long as the process, claiming slices from an atomic counter alongside the caller: the
    // pass itself 0.43 -> 0.25 ms, the call 1.03 -> 1.14 ms (1.49 with three helpers) — a slice claimed by a thread
    // that wakes late, on a core whose cache has never seen the handles, is a slice the caller then waits for.)
    std::vector<char> prepared(n_slices, 0);
    auto ready = [&](size_t sl) {
        if (prepared[sl]) return;
        prepare(sl);
        prepared[sl] = 1;
    };
    const bool force_masks = getenv("VX_STAGE_MASKS") != nullptr, poison = getenv("VX_STAGE_POISON") != nullptr,
               no_list = getenv("VX_STAGE_NO_LIST") != nullptr;
    CU_TRY(cudaEventRecord(it->ev_done[0], s));  // the stage stream starts after whatever is queued (a reset, say)
    CU_TRY(cudaEventRecord(it->ev_done[1], s));
    double host_lists = 0;
    size_t total_listed = 0;
    struct Slice {
        u8 *dm, *dv;
        const u8* k_flags;
        const u64* k_old;
        bool listed;
    };
    std::vector<Slice> sls(n_slices);
    // bus traffic of slice sl: descriptors by the copy engine, then the staging kernel, on the copy stream
    auto enqueue_stage = [&](size_t sl) -> int {
        const double t0 = now_us();
        ready(sl);
        host_lists += now_us() - t0;
        const Prep& P = prep[sl];
        if (P.rc != VX_OK || !P.fused) return VX_OK;  // settled by the caller once every slice has been looked at
        const size_t lo = cut[sl], cnt = cut[sl + 1] - lo, k0 = lo * upc, nu = P.nu;
        total_listed += nu;
        const int sb = int(sl & 1);
        Slice& S = sls[sl];
        S.dm = dbase + size_t(sb) * slab_bytes;
        S.dv = S.dm + up(slice * mbytes);
        CU_TRY(cudaStreamWaitEvent(cs, it->ev_done[sb], 0));  // slab sb is free again
        if (trace) CU_TRY(cudaEventRecord(it->tevs[4 * sl], cs));
        CU_TRY(cudaMemcpyAsync(d_src + lo, h_src + lo, cnt * 8, cudaMemcpyHostToDevice, cs));
        if (nu) CU_TRY(cudaMemcpyAsync(d_units + k0, h_units + k0, nu * 4, cudaMemcpyHostToDevice, cs));
        if (P.any_fill) {  // without a fill the flags add nothing: an untouched batch stages all-zero masks
            CU_TRY(cudaMemcpyAsync(d_flags + lo, h_flags + lo, cnt, cudaMemcpyHostToDevice, cs));
            CU_TRY(cudaMemcpyAsync(d_fills + lo, h_fills + lo, cnt * 8, cudaMemcpyHostToDevice, cs));
        }
        if (P.old_here) CU_TRY(cudaMemcpyAsync(d_old + lo, h_old + lo, cnt * 8, cudaMemcpyHostToDevice, cs));
        S.k_flags = P.any_fill ? d_flags + lo : nullptr;
        S.k_old = P.old_here ? d_old + lo : nullptr;
        // the bulk builder plans from the unit list and never looks at another unit; the fused kernel scans
        // every unit, so for it the slab's masks are zeroed first (in HBM)
        S.listed = bulk_takes_listed(depth, cnt, S.dm, S.k_flags, S.k_old) && !no_list;
        if (poison) {  // tests: nothing outside the staged units / flagged blocks may be consumed
            CU_TRY(cudaMemsetAsync(S.dv, 0xA5, cnt * vbytes, cs));
            CU_TRY(cudaMemsetAsync(S.dm, 0xA5, cnt * mbytes, cs));
        }
        if (!S.listed) CU_TRY(cudaMemsetAsync(S.dm, 0, cnt * mbytes, cs));
        if (nu) {
            // one CTA per SM keeps enough loads on the bus and leaves the SMs to the build of the previous slice
            static const size_t per_sm = getenv("VX_STAGE_CTAS") ? size_t(atoi(getenv("VX_STAGE_CTAS"))) : 1;
            const unsigned grid = unsigned(std::min<size_t>((nu + STAGE_WARPS - 1) / STAGE_WARPS, size_t(it->sm_count) * per_sm));
            // every batch of the slice keeps a journal (API-written, sparse): the packed (block, values) arrays travel
            static const bool no_journal = getenv("VX_STAGE_NO_JOURNAL") != nullptr;
            if (P.journal && !P.raw && !force_masks && !no_journal) {
                CU_TRY(cudaMemcpyAsync(d_jcnt + lo, h_jcnt + lo, cnt * 4, cudaMemcpyHostToDevice, cs));
                if (S.listed)  // (the fused kernel's slab was zeroed whole above)
                    stage_zero_units_kernel<<<grid, STAGE_THREADS, 0, cs>>>(d_units + k0, u32(nu), upc_log2, ub, S.dm, mbytes);
                static const size_t jctas = getenv("VX_STAGE_JCTAS") ? size_t(atoi(getenv("VX_STAGE_JCTAS"))) : 1;  // the bus is the bound, not the SMs
                const unsigned jgrid = unsigned(std::min<size_t>((cnt + STAGE_WARPS - 1) / STAGE_WARPS, size_t(it->sm_count) * jctas));
                if (it->dtype == VX_U8)
                    stage_journal_kernel<8><<<jgrid, STAGE_THREADS, 0, cs>>>(d_src + lo, d_jcnt + lo, u32(cnt), off_jvals, off_jblock,
                                                                            S.dm, S.dv, mbytes, vbytes);
                else
                    stage_journal_kernel<32><<<jgrid, STAGE_THREADS, 0, cs>>>(d_src + lo, d_jcnt + lo, u32(cnt), off_jvals, off_jblock,
                                                                             S.dm, S.dv, mbytes, vbytes);
            } else
            // a batch that only ever went through set/fill/clear/assign has value != 0 <=> set bit, so its masks
            // stay on the host and the one-bit-per-block map travels instead (vx_stage.cuh)
            if (!P.raw && !force_masks) {
                if (it->dtype == VX_U8)
                    stage_units_occ_kernel<8><<<grid, STAGE_THREADS, 0, cs>>>(d_src + lo, d_units + k0, u32(nu), upc_log2, ub,
                                                                            S.dm, S.dv, mbytes, vbytes);
                else
                    stage_units_occ_kernel<32><<<grid, STAGE_THREADS, 0, cs>>>(d_src + lo, d_units + k0, u32(nu), upc_log2, ub,
                                                                             S.dm, S.dv, mbytes, vbytes);
            } else {
                if (it->dtype == VX_U8)
                    stage_units_kernel<8><<<grid, STAGE_THREADS, 0, cs>>>(d_src + lo, d_units + k0, u32(nu), upc_log2, ub, S.dm,
                                                                        S.dv, mbytes, vbytes);
                else
                    stage_units_kernel<32><<<grid, STAGE_THREADS, 0, cs>>>(d_src + lo, d_units + k0, u32(nu), upc_log2, ub, S.dm,
                                                                         S.dv, mbytes, vbytes);
            }
            CU_TRY(cudaGetLastError());
        }
        if (trace) CU_TRY(cudaEventRecord(it->tevs[4 * sl + 1], cs));
        CU_TRY(cudaEventRecord(it->ev_copied[sb], cs));
        return VX_OK;
    };
    auto enqueue_build = [&](size_t sl) -> int {
        const size_t lo = cut[sl], cnt = cut[sl + 1] - lo;
        const int sb = int(sl & 1);
        const Slice& S = sls[sl];
        CU_TRY(cudaStreamWaitEvent(s, it->ev_copied[sb], 0));
        if (trace) CU_TRY(cudaEventRecord(it->tevs[4 * sl + 2], s));
        const bool prof = it->prof;
        it->prof = false;  // the builders' own stage events would reuse the markers
        int r = launch_apply(it, depth, cnt, S.dm, S.dv, S.k_flags, prep[sl].any_fill ? d_fills + lo : nullptr, d_roots + lo,
                             d_changed + lo, s, S.k_old, S.listed ? d_units + lo * upc : nullptr, u32(prep[sl].nu));
        it->prof = prof;
        if (r != VX_OK) return r;
        if (trace) CU_TRY(cudaEventRecord(it->tevs[4 * sl + 3], s));
        CU_TRY(cudaEventRecord(it->ev_done[sb], s));
        return VX_OK;
    };
    // Order of the enqueues: the first two slices' bus traffic goes out before anything else (each into its own
    // slab); then every slice must have been checked; from there on build k is followed by the traffic of k+2,
    // which reuses build k's slab.
    rc = enqueue_stage(0);
    if (rc == VX_OK && n_slices > 1) rc = enqueue_stage(1);
    if (rc != VX_OK) return rc;
    {
        const double t0 = now_us();
        for (size_t sl = 0; sl < n_slices; ++sl) ready(sl);
        host_lists += now_us() - t0;
        for (size_t sl = 0; sl < n_slices; ++sl) {
            fused = fused && prep[sl].fused;
            if (prep[sl].rc != VX_OK) {  // nothing has touched the interner yet
                cudaStreamSynchronize(cs);
                return fail(prep[sl].rc, prep[sl].err);
            }
        }
        if (!fused) {
            cudaStreamSynchronize(cs);
            lk.unlock();
            return serial();
        }
    }
    for (size_t sl = 0; sl < n_slices; ++sl) {
        rc = enqueue_build(sl);
        if (rc == VX_OK && sl + 2 < n_slices) rc = enqueue_stage(sl + 2);
        if (rc != VX_OK) return rc;
    }
    const size_t k = total_listed;
    CU_TRY(cudaMemcpyAsync(h_roots, d_roots, n * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(h_changed, d_changed, n, cudaMemcpyDeviceToHost, s));
    const double t_queued = now_us();
    rc = check_device_error(it);
    if (rc != VX_OK) return rc;
    const double t_synced = now_us();
    if (trace) {
        double stage = 0, build = 0;
        for (size_t sl = 0; sl < n_slices; ++sl) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, it->tevs[4 * sl], it->tevs[4 * sl + 1]);
            cudaEventElapsedTime(&b, it->tevs[4 * sl + 2], it->tevs[4 * sl + 3]);
            stage += a;
            build += b;
        }
        float span = 0;
        cudaEventElapsedTime(&span, it->tevs[0], it->tevs[4 * n_slices - 1]);
        it->trace[0] = host_lists;                          // host: unit lists + descriptors (all slices)
        it->trace[1] = t_queued - t_begin - host_lists;     // host: enqueue
        it->trace[2] = t_synced - t_queued;                 // host: wait for the device
        it->trace[3] = span * 1e3;                          // device: first stage start .. last build end
        it->trace[4] = stage * 1e3;                         // device: descriptors + memset + stage kernel, summed
        it->trace[5] = build * 1e3;                         // device: builds, summed (they overlap the stages)
        it->trace[6] = double(k);
        it->trace[7] = double(n_slices);
    }
    if (it->free_host > 0) {
        rc = refresh_free_count(it);
        if (rc != VX_OK) return rc;
    }
    std::vector<u64> dead;
    for (size_t i = 0; i < n; ++i) {
        // a batch with neither patches nor a fill hands the tree's own root back (voxtree.rs:756-758): on an
        // EMPTY tree apply_batch still answers true and marks it dirty (:303-328).  The device saw such a
        // batch as all-zero masks ("unchanged"), so the reference's answer is put in here.
        if (!(h_flags[i] & (VX_FLAG_FILL | VX_FLAG_PATCHES)) && trees[i]->root == VX_BLOCK_EMPTY) {
            h_changed[i] = 1;
            h_roots[i] = VX_BLOCK_EMPTY;
        }
        if (h_changed[i]) {
            if (trees[i]->root != VX_BLOCK_EMPTY) dead.push_back(trees[i]->root);
            trees[i]->root = h_roots[i];
            trees[i]->dirty = true;
        }
        if (changed) changed[i] = h_changed[i];
    }
    if (!dead.empty()) {  // every build has finished: now drop the old trees (voxtree.rs:309-314)
        rc = release_roots(it, dead.data(), dead.size());
        if (rc != VX_OK) return rc;
    }
    return VX_OK;
}

// ------------------------------------------------------------------------------- read back
int vx_tree_get_many(const vx_interner* cit, const vx_tree* t, size_t n, const int32_t* xyz, uint8_t* found,
                     int64_t* values) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it || !t || !xyz || !found || !values) return fail(VX_E_INVALID, "null argument");
    if (n == 0) return VX_OK;
    const int N = 1 << t->depth;
    for (size_t i = 0; i < 3 * n; ++i)
        if (xyz[i] < 0 || xyz[i] >= N) return fail(VX_E_BOUNDS, "position out of bounds");  // voxtree.rs:146-148
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    int rc = ensure_scratch(it, n * 12 + n + n * 8 + 64, 0);
    if (rc != VX_OK) return rc;
    long long* d_out = (long long*)it->scratch;
    int* d_xyz = (int*)(d_out + n);
    u8* d_found = (u8*)(d_xyz + 3 * n);
    cudaStream_t s = it->stream;
    CU_TRY(cudaMemcpyAsync(d_xyz, xyz, n * 12, cudaMemcpyHostToDevice, s));
    unsigned grid = unsigned((n + 255) / 256);
    if (it->dtype == VX_U8)
        get_many_kernel<u8><<<grid, 256, 0, s>>>(it->dev.children, (const u8*)it->dev.values, t->root, t->depth, n,
                                                  d_xyz, d_found, d_out);
    else
        get_many_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev.children, (const int32_t*)it->dev.values, t->root,
                                                       t->depth, n, d_xyz, d_found, d_out);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(found, d_found, n, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(values, d_out, n * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return VX_OK;
}

int vx_tree_get(const vx_interner* it, const vx_tree* t, int x, int y, int z, int64_t* out) {
    if (!out) return fail(VX_E_INVALID, "null out");
    int32_t xyz[3] = {x, y, z};
    uint8_t found = 0;
    int64_t v = 0;
    int rc = vx_tree_get_many(it, t, 1, xyz, &found, &v);
    if (rc != VX_OK) return rc;
    *out = v;
    return found ? 1 : 0;
}

int vx_roots_to_vec(const vx_interner* cit, uint8_t depth, size_t n, const vx_block_id* roots, void* dense) {
    return vx_roots_to_vec_lod(cit, depth, 0, n, roots, dense);
}

int vx_roots_to_vec_lod(const vx_interner* cit, uint8_t max_depth, uint8_t lod, size_t n, const vx_block_id* roots,
                        void* dense) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it || !roots || !dense) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(max_depth)) return fail(VX_E_INVALID, "max_depth must be in [2,7]");
    if (n == 0) return VX_OK;
    const int depth = max_depth > lod ? int(max_depth) - int(lod) : 0;  // MaxDepth::for_lod saturates
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    const size_t vol = size_t(1) << (3 * depth), esz = dtype_size(it->dtype);
    const bool dev_roots = is_device_ptr(roots), dev_dense = is_device_ptr(dense);
    int rc = ensure_scratch(it, (dev_roots ? 0 : n * 8) + (dev_dense ? 0 : n * vol * esz) + 256, 0);
    if (rc != VX_OK) return rc;
    cudaStream_t s = it->stream;
    u8* p = (u8*)it->scratch;
    const u64* d_roots = roots;
    if (!dev_roots) {
        CU_TRY(cudaMemcpyAsync(p, roots, n * 8, cudaMemcpyHostToDevice, s));
        d_roots = (const u64*)p;
        p += (n * 8 + 255) / 256 * 256;
    }
    void* d_dense = dev_dense ? dense : (void*)p;
    size_t total = n * vol;
    unsigned grid = unsigned(std::min<size_t>((total + 255) / 256, size_t(it->sm_count) * 16));
    if (it->dtype == VX_U8)
        to_vec_kernel<u8><<<grid, 256, 0, s>>>(it->dev.children, (const u8*)it->dev.values, d_roots, n, depth,
                                                (u8*)d_dense);
    else
        to_vec_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev.children, (const int32_t*)it->dev.values, d_roots, n,
                                                     depth, (int32_t*)d_dense);
    CU_TRY(cudaGetLastError());
    if (!dev_dense) CU_TRY(cudaMemcpyAsync(dense, d_dense, total * esz, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return VX_OK;
}

// generate_occupancy_masks — utils/mesh.rs:515-596 for n chunks spread over n_builders OccupancyDataBuilders.
int vx_occupancy_masks(const vx_interner* cit, uint8_t max_depth, uint8_t lod, size_t n, const vx_block_id* roots,
                       const uint32_t* offsets, const uint32_t* builder_of, size_t n_builders, uint32_t max_materials,
                       uint64_t* global, uint64_t* active, uint32_t* n_materials, uint64_t* material_ids,
                       uint64_t* material_counts, uint64_t* per_material) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it || (n && (!roots || !offsets)) || !global || !active || !n_materials || !material_ids ||
        !material_counts || (max_materials && !per_material))
        return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(max_depth)) return fail(VX_E_INVALID, "max_depth must be in [2,7]");
    const int ld = max_depth > lod ? int(max_depth) - int(lod) : 0;  // MaxDepth::for_lod saturates
    if (ld > 6) return fail(VX_E_UNSUPPORTED, "an occupancy volume is 64 voxels per axis (utils/mesh.rs:50): depth - lod <= 6");
    if (n_builders == 0) return n ? fail(VX_E_INVALID, "chunks but no builder") : VX_OK;
    if (max_materials > 1024) return fail(VX_E_INVALID, "max_materials <= 1024");
    if (n_builders > 65535) return fail(VX_E_UNSUPPORTED, "at most 65535 builders per call (one grid row each)");
    const bool dev_out = is_device_ptr(global);
    if (dev_out != is_device_ptr(active) || dev_out != is_device_ptr(n_materials) ||
        dev_out != is_device_ptr(material_ids) || dev_out != is_device_ptr(material_counts) ||
        (max_materials && dev_out != is_device_ptr(per_material)))
        return fail(VX_E_INVALID, "all outputs must live in the same memory space");
    const size_t S = size_t(1) << ld, G = 64 >> ld, cells_per = G * G * G, ncell = n_builders * cells_per;
    // host: place every chunk in its builder's cell grid (offsets are multiples of the chunk side, one chunk per cell)
    std::vector<u64> cells(ncell, 0);
    std::vector<u8> taken(ncell, 0);
    for (size_t i = 0; i < n; ++i) {
        const size_t b = builder_of ? builder_of[i] : 0;
        const uint32_t* o = offsets + 3 * i;
        if (b >= n_builders) return fail(VX_E_INVALID, "builder index out of range");
        if (o[0] % S || o[1] % S || o[2] % S || o[0] + S > 64 || o[1] + S > 64 || o[2] + S > 64)
            return fail(VX_E_BOUNDS, "chunk offset must be a multiple of the chunk side inside the 64^3 volume");
        if (roots[i] == VX_BLOCK_INVALID || id_index(roots[i]) >= it->capacity)
            return fail(VX_E_INVALID, "invalid block id");
        const size_t c = b * cells_per + ((o[1] >> ld) * G + (o[2] >> ld)) * G + (o[0] >> ld);
        if (taken[c]) return fail(VX_E_INVALID, "two chunks at the same offset of one builder");
        taken[c] = 1;
        cells[c] = roots[i];
    }
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    const size_t M = max_materials, plane_bytes = size_t(OCC_ALL) * 8;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t need = up(ncell * 8) + 256 + up(n_builders * 4);
    if (!dev_out)
        need += up(n_builders * plane_bytes) + up(n_builders * 48) + up(n_builders * 4) + 2 * up(n_builders * M * 8) +
                up(n_builders * M * plane_bytes);
    int rc = ensure_scratch(it, need, 0);
    if (rc != VX_OK) return rc;
    cudaStream_t s = it->stream;
    u8* p = (u8*)it->scratch;
    auto take = [&](size_t bytes) { u8* r = p; p += up(bytes); return r; };
    u32* d_err = (u32*)take(256);
    u64* d_cells = (u64*)take(ncell * 8);
    u32* d_over = (u32*)take(n_builders * 4);
    u64 *d_global = global, *d_active = active, *d_ids = material_ids, *d_counts = material_counts, *d_pm = per_material;
    u32* d_nmat = n_materials;
    if (!dev_out) {
        d_global = (u64*)take(n_builders * plane_bytes);
        d_active = (u64*)take(n_builders * 48);
        d_nmat = (u32*)take(n_builders * 4);
        d_ids = (u64*)take(n_builders * M * 8);
        d_counts = (u64*)take(n_builders * M * 8);
        d_pm = (u64*)take(n_builders * M * plane_bytes);
    }
    u32* d_over_count = d_err + 1;
    CU_TRY(cudaMemsetAsync(d_err, 0, 8, s));
    CU_TRY(cudaMemcpyAsync(d_cells, cells.data(), ncell * 8, cudaMemcpyHostToDevice, s));
    // 1. shared-memory path: every builder with <= ms materials (vx_occupancy.cuh: occ_planes_kernel), chunks >= 8^3
    const int ms = int(std::min<uint32_t>(max_materials, OCC_MS_MAX));
    u32 n_over = u32(n_builders);
    const u32* d_only = nullptr;  // word-owner kernels: every builder
    if (ld >= 3) {
        const size_t smem = size_t(std::max(ms, 1)) * OCC_HALVES * 4;
        const dim3 grid_planes(3, unsigned(n_builders));
        // <= 3 materials: 96 KiB of planes -> two CTAs of 512 threads per SM (their barriers overlap); else one of 1024
        const unsigned nthr = ms <= 3 ? 512u : 1024u;
        if (it->dtype == VX_U8) {
            CU_TRY(cudaFuncSetAttribute(occ_planes_kernel<u8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            occ_planes_kernel<u8><<<grid_planes, nthr, smem, s>>>(it->dev.children, (const u8*)it->dev.values, d_cells,
                                                                   ld, ms, max_materials, d_nmat, d_ids, d_counts,
                                                                   d_global, d_active, d_pm, d_over, d_over_count);
        } else {
            CU_TRY(cudaFuncSetAttribute(occ_planes_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            occ_planes_kernel<int32_t><<<grid_planes, nthr, smem, s>>>(it->dev.children, (const int32_t*)it->dev.values,
                                                                        d_cells, ld, ms, max_materials, d_nmat, d_ids,
                                                                        d_counts, d_global, d_active, d_pm, d_over,
                                                                        d_over_count);
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(&n_over, d_over_count, 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        d_only = d_over;
    }
    // 2. builders with more materials than fit in shared memory: word-owner kernels, only for the flagged ones
    if (n_over) {
        const dim3 grid_masks(OCC_ALL / 256, unsigned(n_builders));
        if (it->dtype == VX_U8) {
            occ_materials_kernel<u8><<<unsigned(n_builders), 256, 0, s>>>(it->dev.children, (const u8*)it->dev.values,
                                                                           d_cells, ld, max_materials, d_nmat, d_ids,
                                                                           d_counts, d_active, d_err, d_only);
            occ_masks_kernel<u8><<<grid_masks, 256, 0, s>>>(it->dev.children, (const u8*)it->dev.values, d_cells, ld,
                                                             max_materials, d_nmat, d_ids, d_global, d_active, d_pm,
                                                             d_only);
        } else {
            occ_materials_kernel<int32_t><<<unsigned(n_builders), 256, 0, s>>>(
                it->dev.children, (const int32_t*)it->dev.values, d_cells, ld, max_materials, d_nmat, d_ids, d_counts,
                d_active, d_err, d_only);
            occ_masks_kernel<int32_t><<<grid_masks, 256, 0, s>>>(it->dev.children, (const int32_t*)it->dev.values,
                                                                  d_cells, ld, max_materials, d_nmat, d_ids, d_global,
                                                                  d_active, d_pm, d_only);
        }
        CU_TRY(cudaGetLastError());
    }
    u32 err = 0;
    CU_TRY(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, s));
    if (!dev_out) {
        CU_TRY(cudaMemcpyAsync(n_materials, d_nmat, n_builders * 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(global, d_global, n_builders * plane_bytes, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(active, d_active, n_builders * 48, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(material_ids, d_ids, n_builders * M * 8, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(material_counts, d_counts, n_builders * M * 8, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (err == OCC_ERR_MATERIALS)
        return fail(VX_E_BOUNDS, "a builder holds more materials than max_materials (n_materials[] has the counts)");
    if (!dev_out) {  // only the planes of materials that exist travel back
        for (size_t b = 0; b < n_builders; ++b)
            if (n_materials[b])
                CU_TRY(cudaMemcpyAsync(per_material + b * M * OCC_ALL, d_pm + b * M * OCC_ALL,
                                       size_t(n_materials[b]) * plane_bytes, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
    }
    return VX_OK;
}

int vx_tree_to_vec(const vx_interner* it, const vx_tree* t, void* dense) {
    if (!t) return fail(VX_E_INVALID, "null tree");
    return vx_roots_to_vec(it, t->depth, 1, &t->root, dense);
}

// ------------------------------------------------------------------------------- fill / clear / set_root
int vx_tree_set_root_id(vx_interner* it, vx_tree* t, vx_block_id root) {
    if (!it || !t) return fail(VX_E_INVALID, "null argument");
    if (root == VX_BLOCK_INVALID || id_index(root) >= it->capacity) return fail(VX_E_INVALID, "invalid block id");
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    if (root != VX_BLOCK_EMPTY) {
        // is_valid_block_id (interner/mod.rs:997-1008): the slot must be live and of the id's generation — a stale id
        // would put a reference on a free or recycled slot, which is later freed a second time
        Scalars sc;
        int rc = read_scalars(it, &sc);
        if (rc != VX_OK) return rc;
        const u32 idx = id_index(root);
        u16 gen = 0;
        u64 hash = 0;
        u32 ref = 0;
        if (idx == 0 || idx >= sc.next_index) return fail(VX_E_INVALID, "invalid block id (index was never allocated)");
        CU_TRY(cudaMemcpyAsync(&gen, it->dev.gens + idx, 2, cudaMemcpyDeviceToHost, it->stream));
        CU_TRY(cudaMemcpyAsync(&hash, it->dev.hashes + idx, 8, cudaMemcpyDeviceToHost, it->stream));
        CU_TRY(cudaMemcpyAsync(&ref, it->dev.refs + idx, 4, cudaMemcpyDeviceToHost, it->stream));
        CU_TRY(cudaStreamSynchronize(it->stream));
        if (hash == 0 || ref == 0 || gen != u16((root >> 32) & 0x7FFF))
            return fail(VX_E_INVALID, "invalid block id (slot is free or belongs to another generation)");
    }
    t->root = root;  // voxtree.rs:135-141: adopt + inc_ref (the previous root is NOT released, as in the reference)
    add_ref_kernel<<<1, 1, 0, it->stream>>>(it->dev, root, 1u);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(it->stream));
    return VX_OK;
}

int vx_tree_clear(vx_interner* it, vx_tree* t) {
    if (!it || !t) return fail(VX_E_INVALID, "null argument");
    if (t->root == VX_BLOCK_EMPTY) return VX_OK;  // voxtree.rs:283-292
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    u64 old_root = t->root;
    int rc = release_roots(it, &old_root, 1);
    if (rc != VX_OK) return rc;
    t->root = VX_BLOCK_EMPTY;
    t->dirty = true;
    return VX_OK;
}

int vx_tree_fill(vx_interner* it, vx_tree* t, int64_t value) {
    if (!it || !t) return fail(VX_E_INVALID, "null argument");
    int64_t v = it->dtype == VX_U8 ? int64_t(u8(value)) : int64_t(int32_t(value));
    if (v == 0) return vx_tree_clear(it, t);  // voxtree.rs:269-279
    if (t->root != VX_BLOCK_EMPTY) {          // release the old tree first (:271-273)
        int rc = vx_tree_clear(it, t);
        if (rc != VX_OK) return rc;
    }
    // get_or_create_leaf(value) + root handle == a fill-only batch on the now empty tree
    uint8_t flag = VX_FLAG_FILL;
    vx_block_id root = 0;
    uint8_t changed = 0;
    int rc;
    {
        std::lock_guard<std::mutex> lk(it->mu);
        DeviceGuard g(it->device);
        rc = ensure_scratch(it, 1 << 20, 0);
    }
    if (rc != VX_OK) return rc;
    // masks/values are not read for a fill-only chunk; any valid device pointer will do
    u8* dummy = (u8*)it->scratch + (1 << 19);
    rc = vx_apply_batches_slab(it, t->depth, 1, dummy, dummy, &flag, &v, &root, &changed);
    if (rc != VX_OK) return rc;
    t->root = root;
    t->dirty = true;
    return VX_OK;
}

// ------------------------------------------------------------------------------- global dedup
// Synchronous helpers for the hash-partitioned merge of per-GPU interners; the all-to-all itself is done
// by the caller (NCCL through torch.distributed, see voxelis_b200/dedup.py).  All pointers are device memory.
int vx_dedup_heights(vx_interner* it, uint8_t* d_heights) {
    if (!it || !d_heights) return fail(VX_E_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    Scalars sc{};
    cudaStream_t s = it->stream;
    CU_TRY(cudaMemcpyAsync(&sc, it->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const u32 n = std::min<u32>(sc.next_index, u32(it->capacity));
    int rc = ensure_scratch(it, 64, 0);
    if (rc != VX_OK) return rc;
    u32* d_changed = (u32*)it->scratch;
    const unsigned grid = (n + 255) / 256;
    if (it->dtype == VX_U8)
        heights_init_kernel<u8><<<grid, 256, 0, s>>>(it->dev, n, d_heights);
    else
        heights_init_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev, n, d_heights);
    CU_TRY(cudaGetLastError());
    int sweeps = 0;
    for (; sweeps < 64; ++sweeps) {
        u32 changed = 0;
        CU_TRY(cudaMemsetAsync(d_changed, 0, 4, s));
        heights_sweep_kernel<<<grid, 256, 0, s>>>(it->dev, n, d_heights, d_changed);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(&changed, d_changed, 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        if (!changed) break;
    }
    return sweeps;  // == height of the tallest node
}

int vx_dedup_pack(vx_interner* it, int height, const uint8_t* d_heights, const uint64_t* d_gmap, int G,
                  uint64_t* counts_out, uint64_t* d_records, uint32_t* d_src) {
    if (!it || !d_heights || !d_gmap || !counts_out || G < 1 || G > 8) return fail(VX_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    Scalars sc{};
    cudaStream_t s = it->stream;
    CU_TRY(cudaMemcpyAsync(&sc, it->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const u32 n = std::min<u32>(sc.next_index, u32(it->capacity));
    int rc = ensure_scratch(it, 256, 0);
    if (rc != VX_OK) return rc;
    u32* d_counts = (u32*)it->scratch;   // [8] counts, [8] bases, [8] cursors
    u32* d_bases = d_counts + 8;
    u32* d_cursors = d_counts + 16;
    const unsigned grid = (n + 255) / 256;
    CU_TRY(cudaMemsetAsync(d_counts, 0, 96, s));
    if (it->dtype == VX_U8)
        dedup_count_kernel<u8><<<grid, 256, 0, s>>>(it->dev, n, d_heights, u32(height), d_gmap, u32(G), d_counts);
    else
        dedup_count_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev, n, d_heights, u32(height), d_gmap, u32(G), d_counts);
    CU_TRY(cudaGetLastError());
    u32 h_counts[8] = {0};
    CU_TRY(cudaMemcpyAsync(h_counts, d_counts, 32, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    u32 h_bases[8] = {0};
    for (int o = 0; o < G; ++o) {
        counts_out[o] = h_counts[o];
        if (o) h_bases[o] = h_bases[o - 1] + h_counts[o - 1];
    }
    if (!d_records) return VX_OK;  // counting call
    if (!d_src) return fail(VX_E_INVALID, "null src");
    CU_TRY(cudaMemcpyAsync(d_bases, h_bases, 32, cudaMemcpyHostToDevice, s));
    if (it->dtype == VX_U8)
        dedup_fill_kernel<u8><<<grid, 256, 0, s>>>(it->dev, n, d_heights, u32(height), d_gmap, u32(G), d_bases, d_cursors,
                                                    d_records, d_src);
    else
        dedup_fill_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev, n, d_heights, u32(height), d_gmap, u32(G), d_bases,
                                                         d_cursors, d_records, d_src);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s));
    return VX_OK;
}

int vx_dedup_scatter(vx_interner* it, size_t n, const uint32_t* d_src, const uint64_t* d_ids, uint64_t* d_gmap) {
    if (!it || (n && (!d_src || !d_ids || !d_gmap))) return fail(VX_E_INVALID, "null argument");
    if (n == 0) return VX_OK;
    DeviceGuard g(it->device);
    dedup_scatter_kernel<<<unsigned((n + 255) / 256), 256, 0, it->stream>>>(u32(n), d_src, d_ids, d_gmap);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(it->stream));
    return VX_OK;
}

int vx_dedup_map_roots(vx_interner* it, size_t n, const uint64_t* d_roots, const uint64_t* d_gmap, uint64_t* d_out) {
    if (!it || (n && (!d_roots || !d_gmap || !d_out))) return fail(VX_E_INVALID, "null argument");
    if (n == 0) return VX_OK;
    DeviceGuard g(it->device);
    dedup_map_roots_kernel<<<unsigned((n + 255) / 256), 256, 0, it->stream>>>(u32(n), d_roots, d_gmap, d_out);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(it->stream));
    return VX_OK;
}

int vx_interner_intern_records(vx_interner* shard, size_t n, const uint64_t* d_records, int rank, int leaf_round,
                               uint64_t* d_ids_out, uint64_t* created_out) {
    if (!shard || (n && (!d_records || !d_ids_out))) return fail(VX_E_INVALID, "null argument");
    if (shard->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    if (created_out) *created_out = 0;
    if (n == 0) return VX_OK;
    std::lock_guard<std::mutex> lk(shard->mu);
    DeviceGuard g(shard->device);
    int rc = ensure_scratch(shard, 64, 0);
    if (rc != VX_OK) return rc;
    u32* d_created = (u32*)shard->scratch;
    cudaStream_t s = shard->stream;
    CU_TRY(cudaMemsetAsync(d_created, 0, 4, s));
    const unsigned grid = unsigned((n + 255) / 256);
    if (shard->dtype == VX_U8)
        intern_records_kernel<u8><<<grid, 256, 0, s>>>(shard->dev, u32(n), d_records, u32(rank), leaf_round != 0, d_ids_out,
                                                        d_created);
    else
        intern_records_kernel<int32_t><<<grid, 256, 0, s>>>(shard->dev, u32(n), d_records, u32(rank), leaf_round != 0,
                                                             d_ids_out, d_created);
    CU_TRY(cudaGetLastError());
    u32 created = 0;
    CU_TRY(cudaMemcpyAsync(&created, d_created, 4, cudaMemcpyDeviceToHost, s));
    rc = check_device_error(shard);
    if (rc != VX_OK) return rc;
    if (created_out) *created_out = created;
    return VX_OK;
}

}  // extern "C"

#include "vx_capi_vtm.inl"
#include "vx_world.cuh"

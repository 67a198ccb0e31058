// vx_world.cuh — the multi-GPU side of the C ABI: one `vx_world` per process (= per GPU), and the global-dedup merge
// of the per-GPU interners behind ONE call (BASELINE.json config 5).  Included at the end of vx_capi.cu
// (same translation unit: it uses the interner's internals).
//
// Reference model being matched: ONE interner for every chunk of the world (world/voxmodel.rs:31-32, one shared
// `Arc<RwLock<VoxInterner>>`); the reference's design document proposes sharding that map by `hash % N` once it is
// the bottleneck (Voxelis Bible §3.9, §13).  vx_world_global_dedup is that proposal for N processes x 1 GPU:
//
//   for height h = 0 (leaves) .. H:                   parents need their children's GLOBAL ids, so leaves go first
//     dedup_pack_kernel      every local node of height h -> its owner's send region      (one pass, counts on the device)
//     ncclAllGather          the G x G matrix of record counts                            (one small D2H + sync per round:
//                                                                                          NCCL's send / recv sizes are host values)
//     ncclSend / ncclRecv    72-byte records to their owners, grouped                     (NVLink / NVSwitch)
//     intern_records_kernel  owners intern what arrived into their global shard           (thread per key, vx_dedup.cuh)
//     ncclSend / ncclRecv    8-byte global ids back to the senders
//     dedup_scatter_regions  gmap[local node] = global id
//
// NCCL is looked up at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch already
// loaded), so the library has no link-time dependency on it; without NCCL the world entries fail with
// VX_E_UNSUPPORTED and everything else works.  The unique id travels between the processes by whatever the host
// has (MPI, torch.distributed, a file): vx_world_unique_id on rank 0, vx_world_create everywhere.
#pragma once
#include <nccl.h>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        auto sym = [&](const char* n) { return dlsym(api.lib, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
                 api.Recv && api.AllGather && api.AllReduce && api.GetErrorString;
    });
    return api;
}

#define NC_TRY(expr)                                                                                      \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess)                                                                            \
            return fail(VX_E_CUDA, std::string(#expr) + ": " + nccl_api().GetErrorString(_r));            \
    } while (0)

struct WorldBuf {  // grow-only device buffer
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t want) {
        if (want <= bytes) return VX_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        want = std::max<size_t>(want + want / 4, 1 << 16);
        CU_TRY(cudaMalloc(&p, want));
        bytes = want;
        return VX_OK;
    }
    ~WorldBuf() {
        if (p) cudaFree(p);
    }
};

}  // namespace

struct vx_world {
    int n = 1, rank = 0, device = 0;
    ncclComm_t comm = nullptr;
    WorldBuf heights, gmap, records, src, back, inbox, ids, small, roots;
    u32* h_small = nullptr;  // pinned: counts matrix / histogram / summary
};

namespace vx {

// ---- kernels of the merge (beside those in vx_dedup.cuh) ---------------------------------------------------------
// histogram of node heights (bins 0..8) + the tallest height, after the sweeps
__global__ void dedup_height_hist_kernel(u32 n, const u8* h, u32* hist) {
    __shared__ u32 sh[16];
    if (threadIdx.x < 16) sh[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 v = h[i];
        if (v < 15) atomicAdd(&sh[v], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 15 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
// heights_sweep_kernel without the "changed" word: the host runs a fixed number of sweeps (heights <= 8)
__global__ void dedup_sweep_kernel(InternerDev in, u32 n, u8* h) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0 || h[i] != H_DEAD || in.hashes[i] == 0) return;
    u32 mx = 0;
    for (int k = 0; k < 8; ++k) {
        u64 ch = in.children[size_t(i) * 8 + k];
        if (ch == 0) continue;
        u32 hc = ((volatile u8*)h)[id_index(ch)];
        if (hc == H_DEAD) return;
        mx = max(mx, hc + 1);
    }
    h[i] = u8(mx);
}
// count + place in ONE pass: owner o's records go to records[o * cap ...], positions taken with an atomic per
// record (warp-aggregated per owner).  counts[o] may run past cap: the host checks after the round.
template <class T>
__global__ void dedup_pack_kernel(InternerDev in, u32 n, const u8* h, u32 height, const u64* gmap, u32 G, u32 cap,
                                  u32* counts, u64* records, u32* src) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < n && h[i] == height;
    u64 kids[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 o = 0;
    if (mine) o = u32(global_key_hash<T>(in, i, gmap, height == 0, kids) >> 40) % G;
    const u32 active = __ballot_sync(FULL, mine);
    if (!mine) return;
    const u32 peers = __match_any_sync(active, o);
    const int leader = __ffs(peers) - 1;
    u32 base = 0;
    if ((threadIdx.x & 31) == leader) base = atomicAdd(&counts[o], u32(__popc(peers)));
    base = __shfl_sync(peers, base, leader);
    const u32 pos = base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
    if (pos >= cap) return;  // reported through counts[o] > cap
    u64* r = records + (size_t(o) * cap + pos) * DEDUP_REC_WORDS;
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = kids[k];
    r[8] = sizeof(T) == 1 ? u64(((const u8*)in.values)[i]) : u64(((const u32*)in.values)[i]);
    src[size_t(o) * cap + pos] = i;
}
// gmap[src[o][j]] = back[o][j] for every owner region at once
__global__ void dedup_scatter_regions_kernel(u32 G, u32 cap, const u32* counts, const u32* src, const u64* back, u64* gmap) {
    const u32 o = blockIdx.y;
    const u32 c = min(counts[o], cap);
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < c; j += gridDim.x * blockDim.x)
        gmap[src[size_t(o) * cap + j]] = back[size_t(o) * cap + j];
}

}  // namespace vx

extern "C" {

int vx_world_unique_id(uint8_t out[VX_WORLD_ID_BYTES]) {
    if (!out) return fail(VX_E_INVALID, "null argument");
    NcclApi& nc = nccl_api();
    if (!nc.ok) return fail(VX_E_UNSUPPORTED, "NCCL (libnccl.so.2) not found");
    static_assert(sizeof(ncclUniqueId) == VX_WORLD_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC_TRY(nc.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return VX_OK;
}

vx_world* vx_world_create(int n_ranks, int rank, const uint8_t id_bytes[VX_WORLD_ID_BYTES], int device) {
    NcclApi& nc = nccl_api();
    if (!nc.ok) {
        fail(VX_E_UNSUPPORTED, "NCCL (libnccl.so.2) not found");
        return nullptr;
    }
    if (n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks || !id_bytes) {
        fail(VX_E_INVALID, "1 <= n_ranks <= 8 (the owner field of a global id has 3 bits), 0 <= rank < n_ranks");
        return nullptr;
    }
    DeviceGuard g(device);
    if (!g.ok) {
        fail(VX_E_CUDA, "cudaSetDevice failed");
        return nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    vx_world* w = new vx_world();
    w->n = n_ranks;
    w->rank = rank;
    w->device = device;
    ncclResult_t r = nc.CommInitRank(&w->comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        fail(VX_E_CUDA, std::string("ncclCommInitRank: ") + nc.GetErrorString(r));
        delete w;
        return nullptr;
    }
    if (cudaMallocHost(&w->h_small, 4096) != cudaSuccess) {
        fail(VX_E_CUDA, "cudaMallocHost failed");
        nc.CommDestroy(w->comm);
        delete w;
        return nullptr;
    }
    return w;
}

void vx_world_destroy(vx_world* w) {
    if (!w) return;
    DeviceGuard g(w->device);
    if (w->comm) nccl_api().CommDestroy(w->comm);
    if (w->h_small) cudaFreeHost(w->h_small);
    delete w;
}

int vx_world_size(const vx_world* w) { return w ? w->n : 0; }
int vx_world_rank(const vx_world* w) { return w ? w->rank : -1; }

// Cross-rank barrier on the device (a one-word all-reduce), with no host synchronisation.
int vx_world_barrier(vx_world* w, void* stream) {
    if (!w) return fail(VX_E_INVALID, "null argument");
    DeviceGuard g(w->device);
    int rc = w->small.ensure(4096);
    if (rc != VX_OK) return rc;
    NcclApi& nc = nccl_api();
    NC_TRY(nc.AllReduce(w->small.p, w->small.p, 1, ncclUint32, ncclSum, w->comm, (cudaStream_t)stream));
    return VX_OK;
}

int vx_world_global_dedup(vx_world* w, vx_interner* local, vx_interner* shard, size_t n_roots, const vx_block_id* roots,
                          vx_block_id* global_roots_out, vx_dedup_summary* out) {
    if (!w || !local || !shard || (n_roots && (!roots || !global_roots_out))) return fail(VX_E_INVALID, "null argument");
    if (local->dtype != shard->dtype) return fail(VX_E_INVALID, "local interner and global shard differ in dtype");
    if (local->device != w->device || shard->device != w->device) return fail(VX_E_INVALID, "interners must live on the world's device");
    if (local->poisoned || shard->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    if (local == shard) return fail(VX_E_INVALID, "the global shard must be a separate interner");
    NcclApi& nc = nccl_api();
    std::lock_guard<std::mutex> lk(local->mu);
    std::lock_guard<std::mutex> lk2(shard->mu);
    DeviceGuard g(w->device);
    const u32 G = u32(w->n), me = u32(w->rank);
    cudaStream_t s = local->stream;
    CU_TRY(cudaStreamSynchronize(shard->stream));  // the shard's kernels run on the local interner's stream from here on
    const bool is_u8 = local->dtype == VX_U8;
    // ---- node heights (fixed number of sweeps: a D <= 7 tree has heights <= 7), histogram, tallest height anywhere
    Scalars sc{};
    CU_TRY(cudaMemcpyAsync(&sc, local->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const u32 n = std::min<u32>(sc.next_index, u32(local->capacity));
    int rc;
    if ((rc = w->heights.ensure(n)) != VX_OK || (rc = w->gmap.ensure(size_t(n) * 8)) != VX_OK || (rc = w->small.ensure(4096)) != VX_OK)
        return rc;
    u8* d_h = (u8*)w->heights.p;
    u64* d_gmap = (u64*)w->gmap.p;
    u32* d_small = (u32*)w->small.p;  // [0..15] height histogram, [16..16+G) counts, [32..32+G*G) all counts, [128..] created
    u32* d_hist = d_small;
    u32* d_counts = d_small + 16;
    u32* d_all = d_small + 32;
    u32* d_created = d_small + 128;  // [0] branches, [1] leaves
    CU_TRY(cudaMemsetAsync(d_small, 0, 1024, s));
    CU_TRY(cudaMemsetAsync(d_gmap, 0, size_t(n) * 8, s));
    const unsigned grid = (n + 255) / 256;
    if (is_u8)
        heights_init_kernel<u8><<<grid, 256, 0, s>>>(local->dev, n, d_h);
    else
        heights_init_kernel<int32_t><<<grid, 256, 0, s>>>(local->dev, n, d_h);
    for (int sweep = 0; sweep < 8; ++sweep) dedup_sweep_kernel<<<grid, 256, 0, s>>>(local->dev, n, d_h);
    dedup_height_hist_kernel<<<std::min(grid, 1024u), 256, 0, s>>>(n, d_h, d_hist);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(w->h_small, d_hist, 64, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    u32 max_h = 0, max_level = 1;
    for (u32 h = 0; h < 15; ++h)
        if (w->h_small[h]) {
            max_h = h;
            max_level = std::max(max_level, w->h_small[h]);
        }
    {   // tallest height over all ranks (everyone runs the same number of rounds)
        u32* d_mh = d_small + 160;
        CU_TRY(cudaMemcpyAsync(d_mh, &max_h, 4, cudaMemcpyHostToDevice, s));
        NC_TRY(nc.AllReduce(d_mh, d_mh, 1, ncclUint32, ncclMax, w->comm, s));
        CU_TRY(cudaMemcpyAsync(w->h_small, d_mh, 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        max_h = w->h_small[0];
    }
    // per-owner send regions: the hash spreads a level's nodes evenly; twice the mean plus slack, at most the level
    const u32 cap = G == 1 ? max_level : std::min<u32>(max_level, max_level / G * 2 + 4096);
    if ((rc = w->records.ensure(size_t(G) * cap * DEDUP_REC_WORDS * 8)) != VX_OK || (rc = w->src.ensure(size_t(G) * cap * 4)) != VX_OK ||
        (rc = w->back.ensure(size_t(G) * cap * 8)) != VX_OK)
        return rc;
    u64* d_rec = (u64*)w->records.p;
    u32* d_src = (u32*)w->src.p;
    u64* d_back = (u64*)w->back.p;
    uint64_t bytes_sent = 0;
    for (u32 h = 0; h <= max_h; ++h) {
        CU_TRY(cudaMemsetAsync(d_counts, 0, 32, s));
        if (is_u8)
            dedup_pack_kernel<u8><<<grid, 256, 0, s>>>(local->dev, n, d_h, h, d_gmap, G, cap, d_counts, d_rec, d_src);
        else
            dedup_pack_kernel<int32_t><<<grid, 256, 0, s>>>(local->dev, n, d_h, h, d_gmap, G, cap, d_counts, d_rec, d_src);
        CU_TRY(cudaGetLastError());
        NC_TRY(nc.AllGather(d_counts, d_all, 8, ncclUint32, w->comm, s));   // row r = what rank r sends to each owner
        CU_TRY(cudaMemcpyAsync(w->h_small, d_all, size_t(G) * 32, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        const u32* all = w->h_small;
        size_t n_in = 0;
        for (u32 r = 0; r < G; ++r) {
            for (u32 o = 0; o < G; ++o)
                if (all[r * 8 + o] > cap && r == me) return fail(VX_E_CUDA, "global dedup: an owner's send region overflowed");
            n_in += all[r * 8 + me];
        }
        if ((rc = w->inbox.ensure(std::max<size_t>(n_in, 1) * DEDUP_REC_WORDS * 8)) != VX_OK || (rc = w->ids.ensure(std::max<size_t>(n_in, 1) * 8)) != VX_OK)
            return rc;
        u64* d_in = (u64*)w->inbox.p;
        u64* d_ids = (u64*)w->ids.p;
        // records to their owners
        NC_TRY(nc.GroupStart());
        size_t off = 0;
        for (u32 p = 0; p < G; ++p) {
            const u32 sc_ = all[me * 8 + p], rc_ = all[p * 8 + me];
            if (sc_) NC_TRY(nc.Send(d_rec + size_t(p) * cap * DEDUP_REC_WORDS, size_t(sc_) * DEDUP_REC_WORDS, ncclUint64, int(p), w->comm, s));
            if (rc_) NC_TRY(nc.Recv(d_in + off * DEDUP_REC_WORDS, size_t(rc_) * DEDUP_REC_WORDS, ncclUint64, int(p), w->comm, s));
            off += rc_;
            bytes_sent += uint64_t(sc_) * DEDUP_REC_WORDS * 8;
        }
        NC_TRY(nc.GroupEnd());
        if (n_in) {
            const unsigned g2 = unsigned((n_in + 255) / 256);
            if (is_u8)
                intern_records_kernel<u8><<<g2, 256, 0, s>>>(shard->dev, u32(n_in), d_in, me, h == 0, d_ids, d_created + (h == 0 ? 1 : 0));
            else
                intern_records_kernel<int32_t><<<g2, 256, 0, s>>>(shard->dev, u32(n_in), d_in, me, h == 0, d_ids, d_created + (h == 0 ? 1 : 0));
            CU_TRY(cudaGetLastError());
        }
        // global ids back to the senders
        NC_TRY(nc.GroupStart());
        off = 0;
        for (u32 p = 0; p < G; ++p) {
            const u32 sc_ = all[me * 8 + p], rc_ = all[p * 8 + me];
            if (rc_) NC_TRY(nc.Send(d_ids + off, rc_, ncclUint64, int(p), w->comm, s));
            if (sc_) NC_TRY(nc.Recv(d_back + size_t(p) * cap, sc_, ncclUint64, int(p), w->comm, s));
            off += rc_;
            bytes_sent += uint64_t(rc_) * 8;
        }
        NC_TRY(nc.GroupEnd());
        dedup_scatter_regions_kernel<<<dim3(std::max(1u, std::min((cap + 255) / 256, 256u)), G), 256, 0, s>>>(G, cap, d_counts, d_src, d_back, d_gmap);
        CU_TRY(cudaGetLastError());
    }
    // ---- the chunks' roots in global ids; totals over all ranks
    if (n_roots) {
        if ((rc = w->roots.ensure(n_roots * 16)) != VX_OK) return rc;
        u64* d_r = (u64*)w->roots.p;
        CU_TRY(cudaMemcpyAsync(d_r, roots, n_roots * 8, cudaMemcpyHostToDevice, s));
        dedup_map_roots_kernel<<<unsigned((n_roots + 255) / 256), 256, 0, s>>>(u32(n_roots), d_r, d_gmap, d_r + n_roots);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(global_roots_out, d_r + n_roots, n_roots * 8, cudaMemcpyDeviceToHost, s));
    }
    unsigned long long* d_tot = (unsigned long long*)(d_small + 192);  // branches, leaves, bytes, local nodes
    CU_TRY(cudaMemcpyAsync(w->h_small, d_created, 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const u32 my_b = w->h_small[0], my_l = w->h_small[1];
    unsigned long long tot[4] = {my_b, my_l, bytes_sent, n ? n - 1ull : 0ull};
    CU_TRY(cudaMemcpyAsync(d_tot, tot, sizeof(tot), cudaMemcpyHostToDevice, s));
    NC_TRY(nc.AllReduce(d_tot, d_tot, 4, ncclUint64, ncclSum, w->comm, s));
    CU_TRY(cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    rc = check_device_error(shard);
    if (rc != VX_OK) return rc;
    if (out) {
        out->rounds = max_h + 1;
        out->branches = tot[0];
        out->leaves = tot[1];
        out->bytes_sent = tot[2];
        out->local_nodes_all_ranks = tot[3];
        out->this_shard_branches = my_b;
        out->this_shard_leaves = my_l;
    }
    return VX_OK;
}

}  // extern "C"

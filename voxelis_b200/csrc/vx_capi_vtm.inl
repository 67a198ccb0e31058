// vx_capi_vtm.inl — host side of the VTM writer / reader (world/voxmodel.rs:177-408, world/voxchunk.rs:382-440, io/export.rs:90-151,
// io/import.rs:14-98, io/varint.rs): MD5, varints, payload assembly, parsing + validation, zstd through dlopen.  Textually
// included by vx_capi.cu right after its extern "C" block closes (the section opens its own).
// ------------------------------------------------------------------------------- VTM (world/voxmodel.rs, io/export.rs)
namespace {

struct DevBuf {  // scoped device allocation
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

// MD5 (RFC 1321) of the payload: io/export.rs:124-128 stores it ahead of the (compressed) data.
void md5_digest(const u8* data, size_t len, u8 out[16]) {
    static const u32 K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
        0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
        0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
        0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
        0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
        0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const u8 R[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                             4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    u32 h[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
    std::vector<u8> msg(data, data + len);
    msg.push_back(0x80);
    while (msg.size() % 64 != 56) msg.push_back(0);
    const u64 bits = u64(len) * 8;
    for (int i = 0; i < 8; ++i) msg.push_back(u8(bits >> (8 * i)));
    for (size_t off = 0; off < msg.size(); off += 64) {
        u32 w[16];
        for (int i = 0; i < 16; ++i)
            w[i] = u32(msg[off + 4 * i]) | u32(msg[off + 4 * i + 1]) << 8 | u32(msg[off + 4 * i + 2]) << 16 | u32(msg[off + 4 * i + 3]) << 24;
        u32 a = h[0], b = h[1], c = h[2], d = h[3];
        for (int i = 0; i < 64; ++i) {
            u32 f, g;
            if (i < 16)
                f = (b & c) | (~b & d), g = u32(i);
            else if (i < 32)
                f = (d & b) | (~d & c), g = u32(5 * i + 1) & 15;
            else if (i < 48)
                f = b ^ c ^ d, g = u32(3 * i + 5) & 15;
            else
                f = c ^ (b | ~d), g = u32(7 * i) & 15;
            const u32 t = d;
            d = c;
            c = b;
            const u32 x = a + f + K[i] + w[g];
            b = b + ((x << R[i]) | (x >> (32 - R[i])));
            a = t;
        }
        h[0] += a, h[1] += b, h[2] += c, h[3] += d;
    }
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 4; ++k) out[4 * i + k] = u8(h[i] >> (8 * k));
}

void be32(std::vector<u8>& o, u32 v) {
    for (int sft = 24; sft >= 0; sft -= 8) o.push_back(u8(v >> sft));
}
void host_varint(std::vector<u8>& o, u64 v) {
    while (v >= 0x80) {
        o.push_back(u8((v & 0x7F) | 0x80));
        v >>= 7;
    }
    o.push_back(u8(v));
}

// size_only: *size_out = payload bytes, nothing is written or copied back.
int model_serialize_impl(vx_interner* it, size_t n, const int32_t* positions, const vx_block_id* roots, std::vector<u8>& payload,
                         bool size_only = false, size_t* size_out = nullptr) {
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    cudaStream_t s = it->stream;
    Scalars sc;
    int rc = read_scalars(it, &sc);
    if (rc != VX_OK) return rc;
    const u32 nn = sc.next_index;
    // record sizes and offsets are scanned as u32 (the format's own size field is a u32, io/export.rs:141-143): refuse
    // up front what could wrap — a record is at most varint(5) + mask(1) + 8 child varints(40) + value(4) bytes
    if (size_t(nn) * 50 + size_t(n) * 29 + 12 > 0xFFFFFFFFull) return fail(VX_E_INVALID, "VTM data could exceed 4 GiB");
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t words = up(size_t(nn + 1) * 4);
    size_t scan_tmp = 0;
    CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const u32*)nullptr, (u32*)nullptr, int(nn + 1), s));
    // work arrays + output live in the interner's scratch (grown once, kept): no allocation per export
    const size_t work_bytes = 7 * words + up(scan_tmp) + up(n * 8) + up(n * 4);
    const size_t out_max = up(8 + size_t(nn) * (5 + 1 + 8 * 5 + 4));  // every record at its longest
    rc = ensure_scratch(it, work_bytes + out_max, 0);
    if (rc != VX_OK) return rc;
    u8* base = (u8*)it->scratch;
    VtmArgs a{};
    a.children = it->dev.children;
    a.values = it->dev.values;
    a.refs = it->dev.refs;
    a.n = nn;
    a.vsize = u32(dtype_size(it->dtype));
    a.leaf_flag = (u32*)(base + 0 * words);
    a.branch_flag = (u32*)(base + 1 * words);
    a.leaf_rank = (u32*)(base + 2 * words);
    a.branch_rank = (u32*)(base + 3 * words);
    a.newid = (u32*)(base + 4 * words);
    a.sizes = (u32*)(base + 5 * words);
    a.offs = (u32*)(base + 6 * words);
    void* tmp = base + 7 * words;
    u64* d_roots = (u64*)(base + 7 * words + up(scan_tmp));
    u32* d_root_ids = (u32*)((u8*)d_roots + up(n * 8));
    const unsigned grid = (nn + 255) / 256;
    CU_TRY(cudaMemsetAsync(a.sizes, 0, size_t(nn + 1) * 4, s));
    vtm_classify_kernel<<<grid, 256, 0, s>>>(a);
    CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, a.leaf_flag, a.leaf_rank, int(nn), s));
    CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, a.branch_flag, a.branch_rank, int(nn), s));
    vtm_newid_kernel<<<grid, 256, 0, s>>>(a);
    vtm_sizes_kernel<<<grid, 256, 0, s>>>(a);
    CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, a.sizes, a.offs, int(nn + 1), s));
    CU_TRY(cudaGetLastError());
    // counts: leaves, branches, bytes of the leaf records, bytes of all records
    u32 last[4] = {0, 0, 0, 0};
    CU_TRY(cudaMemcpyAsync(&last[0], a.leaf_rank + nn - 1, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&last[1], a.leaf_flag + nn - 1, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&last[2], a.branch_rank + nn - 1, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&last[3], a.branch_flag + nn - 1, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const u32 n_leaves = last[0] + last[1], n_branches = last[2] + last[3];
    u32 leaf_bytes = 0, node_bytes = 0;
    CU_TRY(cudaMemcpyAsync(&leaf_bytes, a.offs + n_leaves, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&node_bytes, a.offs + n_leaves + n_branches, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    std::vector<u32> root_ids(n);
    if (size_only) {  // the chunk table's size needs the roots' new ids (varints)
        if (n) {
            CU_TRY(cudaMemcpyAsync(d_roots, roots, n * 8, cudaMemcpyHostToDevice, s));
            vtm_roots_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(a.newid, d_roots, u32(n), d_root_ids);
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(root_ids.data(), d_root_ids, n * 4, cudaMemcpyDeviceToHost, s));
            CU_TRY(cudaStreamSynchronize(s));
        }
        size_t total = size_t(node_bytes) + 8 + 4;
        for (size_t c = 0; c < n; ++c) total += 24 + (root_ids[c] < (1u << 7) ? 1 : root_ids[c] < (1u << 14) ? 2 : root_ids[c] < (1u << 21) ? 3 : root_ids[c] < (1u << 28) ? 4 : 5);
        *size_out = total;
        return VX_OK;
    }
    a.out = base + work_bytes;
    vtm_write_kernel<<<grid, 256, 0, s>>>(a);
    CU_TRY(cudaGetLastError());
    if (n) {
        CU_TRY(cudaMemcpyAsync(d_roots, roots, n * 8, cudaMemcpyHostToDevice, s));
        vtm_roots_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(a.newid, d_roots, u32(n), d_root_ids);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(root_ids.data(), d_root_ids, n * 4, cudaMemcpyDeviceToHost, s));
    }
    payload.reserve(size_t(node_bytes) + 12 + n * 29);
    payload.assign(size_t(node_bytes) + 8, 0);
    CU_TRY(cudaMemcpyAsync(payload.data(), a.out, payload.size(), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    auto poke_be32 = [&](size_t at, u32 v) {
        for (int k = 0; k < 4; ++k) payload[at + k] = u8(v >> (24 - 8 * k));
    };
    poke_be32(0, n_leaves);                       // voxmodel.rs:231
    poke_be32(4 + size_t(leaf_bytes), n_branches);  // :242 (the reference's count includes slot 0 and writes count - 1)
    be32(payload, u32(n));                        // :281-284
    static const char VTC_MAGIC[] = "VoxTreeChunk";  // io/consts.rs:3
    for (size_t c = 0; c < n; ++c) {              // world/voxchunk.rs:382-405
        payload.insert(payload.end(), VTC_MAGIC, VTC_MAGIC + 12);
        for (int k = 0; k < 3; ++k) be32(payload, u32(positions[3 * c + k]));
        host_varint(payload, root_ids[c]);
    }
    return VX_OK;
}


struct VtmCursor {
    const u8* p;
    const u8* end;
    bool ok = true;
    u8 byte() {
        if (p >= end) {
            ok = false;
            return 0;
        }
        return *p++;
    }
    u32 be() {
        u32 v = 0;
        for (int i = 0; i < 4; ++i) v = (v << 8) | byte();
        return v;
    }
    u32 varint() {  // io/varint.rs:56-77
        u32 r = 0;
        for (int shift = 0; shift < 35; shift += 7) {
            const u8 b = byte();
            r |= u32(b & 0x7F) << shift;
            if (!(b & 0x80)) return r;
        }
        ok = false;
        return 0;
    }
    int64_t value(size_t bytes) {
        u32 v = 0;
        for (size_t i = 0; i < bytes; ++i) v = (v << 8) | byte();
        return bytes == 1 ? int64_t(v) : int64_t(int32_t(v));
    }
};

int model_deserialize_impl(vx_interner* it, const u8* data, size_t len, int32_t* positions_out, vx_block_id* roots_out, size_t cap,
                           int64_t* n_out) {
    const size_t vs = dtype_size(it->dtype);
    VtmCursor r{data, data + len};
    // ---- parse (voxmodel.rs:310-364, 410; voxchunk.rs:407-440).  File ids must be 1..L for the leaves and
    // L+1.. for the branches, in order: the reference asserts id == next pool index (mod.rs:933,948).
    const u32 L = r.be();
    if (!r.ok || size_t(L) > len) return fail(VX_E_INVALID, "VTM payload: bad leaf count");
    std::vector<u8> values;  // pool image from index 1: [L + Bc] values in device layout
    values.reserve((size_t(L) + 16) * vs);
    auto push_value = [&](int64_t v) {
        if (vs == 1)
            values.push_back(u8(v));
        else {
            const int32_t w = int32_t(v);
            values.insert(values.end(), (const u8*)&w, (const u8*)&w + 4);
        }
    };
    for (u32 k = 0; k < L; ++k) {
        if (r.varint() != k + 1) return fail(VX_E_INVALID, "VTM payload: Invalid block id");
        push_value(r.value(vs));
    }
    const u32 Bc = r.be();
    if (!r.ok || size_t(Bc) > len) return fail(VX_E_INVALID, "VTM payload: bad branch count");
    const size_t N = size_t(L) + Bc;
    if (N + 1 > it->capacity) return fail(VX_E_OOM, "Out of memory");
    std::vector<u32> kids(size_t(Bc) * 8, 0);  // file ids of the children
    std::vector<u8> masks(Bc), types(Bc);
    for (u32 k = 0; k < Bc; ++k) {
        if (r.varint() != L + k + 1) return fail(VX_E_INVALID, "VTM payload: Invalid block id");
        const u8 m = r.byte();
        if (m == 0) return fail(VX_E_INVALID, "VTM payload: branch without children");  // voxmodel.rs:363
        u8 t = 0;
        for (int c = 0; c < 8; ++c) {
            if (!(m >> c & 1)) continue;
            const u32 id = r.varint();
            if (id == 0 || id > N) return fail(VX_E_INVALID, "VTM payload: unknown child id");
            kids[size_t(k) * 8 + c] = id;
            if (id <= L) t |= u8(1u << c);  // leaf_patterns.contains_key (:352-354)
        }
        masks[k] = m;
        types[k] = t;
        push_value(r.value(vs));
    }
    const u32 n = r.be();
    if (!r.ok) return fail(VX_E_INVALID, "VTM payload: truncated");
    if (size_t(n) > cap) return fail(VX_E_INVALID, "vx_model_deserialize: caller arrays too small");
    auto block_of = [&](u32 id) -> u64 {
        if (id == 0) return 0;
        return id <= L ? id_leaf(id) : id_branch(id, types[id - L - 1], masks[id - L - 1]);
    };
    std::vector<u64> rows(size_t(Bc) * 8);
    std::vector<u32> refs(N + 1, 0);
    for (size_t k = 0; k < size_t(Bc) * 8; ++k) {
        rows[k] = block_of(kids[k]);
        if (kids[k]) refs[kids[k]] += 1;  // inc_all_child_refs (mod.rs:999)
    }
    std::vector<u64> roots(n);
    for (u32 c = 0; c < n; ++c) {
        for (int k = 0; k < 12; ++k)
            if (r.byte() != u8("VoxTreeChunk"[k])) return fail(VX_E_INVALID, "VTM payload: bad chunk magic");
        for (int a = 0; a < 3; ++a) {
            const u32 v = r.be();
            if (positions_out) positions_out[3 * size_t(c) + a] = int32_t(v);
        }
        const u32 id = r.varint();
        if (!r.ok || id > N) return fail(VX_E_INVALID, "VTM payload: unknown root id");
        roots[c] = block_of(id);
        if (id) refs[id] += 1;  // set_root_id (voxtree.rs:135-141)
    }
    if (!r.ok) return fail(VX_E_INVALID, "VTM payload: truncated");
    // ---- the graph must be a DAG of DISTINCT nodes before anything touches the interner.  The reference trusts the
    // file (a missing map entry panics, voxmodel.rs:352-360); a crafted or damaged payload could otherwise install a
    // branch that reaches itself (walks never end, refcounts never reach zero) or two nodes with one key (the
    // install kernel's table entries would race and canonicity is lost).  The MD5 only guards against accidents.
    {
        std::unordered_map<int64_t, u32> leaf_seen;
        leaf_seen.reserve(size_t(L) * 2);
        for (u32 k = 0; k < L; ++k) {
            int64_t v = vs == 1 ? int64_t(values[k]) : int64_t(*reinterpret_cast<const int32_t*>(&values[size_t(k) * 4]));
            if (!leaf_seen.emplace(v, k).second) return fail(VX_E_INVALID, "VTM payload: two leaves with the same value");
        }
        // heights by repeated relaxation in file order would be quadratic on a hostile file: iterative DFS instead
        std::vector<u8> state(size_t(Bc), 0);  // 0 = unvisited, 1 = on the stack, 2 = done
        std::vector<std::pair<u32, u8>> stack;
        for (u32 b0 = 0; b0 < Bc; ++b0) {
            if (state[b0]) continue;
            stack.push_back({b0, 0});
            state[b0] = 1;
            while (!stack.empty()) {
                auto& top = stack.back();
                if (top.second == 8) {
                    state[top.first] = 2;
                    stack.pop_back();
                    continue;
                }
                const u32 id = kids[size_t(top.first) * 8 + top.second++];
                if (id <= L) continue;  // EMPTY or a leaf
                const u32 cb = id - L - 1;
                if (state[cb] == 1) return fail(VX_E_INVALID, "VTM payload: a branch reaches itself (cycle)");
                if (state[cb] == 0) {
                    state[cb] = 1;
                    stack.push_back({cb, 0});
                }
            }
        }
        struct RowHash {
            size_t operator()(const std::array<u32, 8>& a) const {
                u64 h = 0;
                for (int i = 0; i < 8; ++i) h = mix64(h + a[i] + 0x9E3779B97F4A7C15ull * (i + 1));
                return size_t(h);
            }
        };
        std::unordered_map<std::array<u32, 8>, u32, RowHash> row_seen;
        row_seen.reserve(size_t(Bc) * 2);
        for (u32 k = 0; k < Bc; ++k) {
            std::array<u32, 8> row;
            for (int c = 0; c < 8; ++c) row[c] = kids[size_t(k) * 8 + c];
            if (!row_seen.emplace(row, k).second) return fail(VX_E_INVALID, "VTM payload: two branches with the same children");
        }
    }
    // ---- install
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    cudaStream_t s = it->stream;
    Scalars sc;
    int rc = read_scalars(it, &sc);
    if (rc != VX_OK) return rc;
    if (sc.next_index != 1 || sc.free_count != 0)
        return fail(VX_E_INVALID, "vx_model_deserialize needs a fresh interner (the reference asserts file id == pool index)");
    if (N) {
        CU_TRY(cudaMemcpyAsync((u8*)it->dev.values + vs, values.data(), N * vs, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(it->dev.refs + 1, refs.data() + 1, N * 4, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemsetAsync(it->dev.children + 8, 0, size_t(L) * 64, s));
        if (Bc) CU_TRY(cudaMemcpyAsync(it->dev.children + (size_t(L) + 1) * 8, rows.data(), size_t(Bc) * 64, cudaMemcpyHostToDevice, s));
        const unsigned grid = unsigned((N + 255) / 256);
        if (it->dtype == VX_U8)
            vtm_install_kernel<u8><<<grid, 256, 0, s>>>(it->dev, L, u32(N));
        else
            vtm_install_kernel<int32_t><<<grid, 256, 0, s>>>(it->dev, L, u32(N));
        CU_TRY(cudaGetLastError());
        const u32 next = u32(N + 1);
        CU_TRY(cudaMemcpyAsync(&it->d_scalars->next_index, &next, 4, cudaMemcpyHostToDevice, s));
    }
    rc = check_device_error(it);
    if (rc != VX_OK) return rc;
    if (roots_out) memcpy(roots_out, roots.data(), size_t(n) * 8);
    *n_out = int64_t(n);
    return VX_OK;
}

// zstd stream -> bytes through libzstd looked up at run time (the reference writes with the streaming encoder,
// io/export.rs:132-136, so the frame need not carry its content size)
int zstd_decompress(const u8* src, size_t len, std::vector<u8>& out) {
    struct InBuf { const void* src; size_t size, pos; };
    struct OutBuf { void* dst; size_t size, pos; };
    typedef void* (*create_fn)();
    typedef size_t (*free_fn)(void*);
    typedef size_t (*step_fn)(void*, OutBuf*, InBuf*);
    typedef unsigned (*err_fn)(size_t);
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
    create_fn create = h ? (create_fn)dlsym(h, "ZSTD_createDStream") : nullptr;
    free_fn destroy = h ? (free_fn)dlsym(h, "ZSTD_freeDStream") : nullptr;
    step_fn step = h ? (step_fn)dlsym(h, "ZSTD_decompressStream") : nullptr;
    err_fn is_err = h ? (err_fn)dlsym(h, "ZSTD_isError") : nullptr;
    if (!create || !destroy || !step || !is_err) return fail(VX_E_UNSUPPORTED, "libzstd not found: cannot read a compressed VTM file");
    void* ds = create();
    if (!ds) return fail(VX_E_INVALID, "ZSTD_createDStream failed");
    InBuf in{src, len, 0};
    std::vector<u8> chunk(size_t(1) << 20);
    size_t hint = 1;
    while (in.pos < in.size || hint != 0) {
        OutBuf ob{chunk.data(), chunk.size(), 0};
        hint = step(ds, &ob, &in);
        if (is_err(hint)) {
            destroy(ds);
            return fail(VX_E_INVALID, "corrupt zstd stream in VTM file");
        }
        out.insert(out.end(), chunk.data(), chunk.data() + ob.pos);
        if (in.pos == in.size && ob.pos == 0) break;
    }
    destroy(ds);
    return VX_OK;
}

}  // namespace

extern "C" {

int64_t vx_model_serialize(const vx_interner* cit, size_t n, const int32_t* positions, const vx_block_id* roots, uint8_t* out,
                           size_t cap) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it || (n && (!positions || !roots))) return fail(VX_E_INVALID, "null argument");
    if (n > 0xFFFFFFFFull) return fail(VX_E_INVALID, "too many chunks");
    std::vector<u8> payload;
    if (!out) {  // size query
        size_t total = 0;
        int rc = model_serialize_impl(it, n, positions, roots, payload, true, &total);
        return rc != VX_OK ? rc : int64_t(total);
    }
    int rc = model_serialize_impl(it, n, positions, roots, payload);
    if (rc != VX_OK) return rc;
    if (payload.size() <= cap) memcpy(out, payload.data(), payload.size());
    return int64_t(payload.size());
}

int vx_export_vtm(const vx_interner* cit, const char* path, const char* name, uint8_t max_depth, float chunk_world_size,
                  const int32_t world_bounds[3], size_t n, const int32_t* positions, const vx_block_id* roots, int compress) {
    vx_interner* it = const_cast<vx_interner*>(cit);
    if (!it || !path || !name || !world_bounds || (n && (!positions || !roots))) return fail(VX_E_INVALID, "null argument");
    if (strlen(name) > 255) return fail(VX_E_INVALID, "model name longer than 255 bytes");  // export.rs:121 (u8 length)
    std::vector<u8> payload;
    int rc = model_serialize_impl(it, n, positions, roots, payload);
    if (rc != VX_OK) return rc;
    u8 digest[16];
    md5_digest(payload.data(), payload.size(), digest);  // export.rs:124-128: over the UNcompressed payload
    std::vector<u8> packed;
    bool compressed = false;
    if (compress) {  // Flags::DEFAULT = COMPRESSED, zstd level 7 (export.rs:132-136); libzstd is looked up at run time
        typedef size_t (*bound_fn)(size_t);
        typedef size_t (*comp_fn)(void*, size_t, const void*, size_t, int);
        typedef unsigned (*err_fn)(size_t);
        void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
        bound_fn bound = h ? (bound_fn)dlsym(h, "ZSTD_compressBound") : nullptr;
        comp_fn comp = h ? (comp_fn)dlsym(h, "ZSTD_compress") : nullptr;
        err_fn is_err = h ? (err_fn)dlsym(h, "ZSTD_isError") : nullptr;
        if (!bound || !comp || !is_err) return fail(VX_E_UNSUPPORTED, "libzstd not found: export with compress = 0 (Flags::NONE)");
        packed.resize(bound(payload.size()));
        const size_t got = comp(packed.data(), packed.size(), payload.data(), payload.size(), 7);
        if (is_err(got)) return fail(VX_E_INVALID, "zstd compression failed");
        packed.resize(got);
        compressed = true;
    }
    const std::vector<u8>& data = compressed ? packed : payload;
    if (data.size() > 0xFFFFFFFFull) return fail(VX_E_INVALID, "VTM data larger than 4 GiB");
    std::vector<u8> head;
    static const char VTM_MAGIC[] = "VoxTreeModel";  // io/consts.rs:1-2
    head.insert(head.end(), VTM_MAGIC, VTM_MAGIC + 12);
    head.push_back(0x01), head.push_back(0x00);             // VTM_VERSION 0x0100, big endian
    head.push_back(0), head.push_back(compressed ? 1 : 0);  // Flags (io/flags.rs)
    head.push_back(max_depth);                              // export.rs:111
    u32 fbits;
    memcpy(&fbits, &chunk_world_size, 4);
    be32(head, fbits);                                      // :112-114
    be32(head, 0), be32(head, 0);                           // RESERVED_1 / RESERVED_2
    for (int k = 0; k < 3; ++k) be32(head, u32(world_bounds[k]));
    head.push_back(u8(strlen(name)));
    head.insert(head.end(), name, name + strlen(name));
    head.insert(head.end(), digest, digest + 16);
    be32(head, u32(data.size()));
    FILE* f = fopen(path, "wb");
    if (!f) return fail(VX_E_INVALID, std::string("cannot open ") + path);
    const bool ok = fwrite(head.data(), 1, head.size(), f) == head.size() && fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    return ok ? VX_OK : fail(VX_E_INVALID, std::string("short write to ") + path);
}

int64_t vx_model_deserialize(vx_interner* it, const uint8_t* data, size_t len, int32_t* positions_out, vx_block_id* roots_out,
                             size_t cap) {
    if (!it || !data) return fail(VX_E_INVALID, "null argument");
    if (it->poisoned) return fail(VX_E_POISONED, "interner overflowed earlier; reset it");
    int64_t n = 0;
    int rc = model_deserialize_impl(it, data, len, positions_out, roots_out, cap, &n);
    return rc != VX_OK ? rc : n;
}

int64_t vx_import_vtm(vx_interner* it, const char* path, vx_vtm_info* info, int32_t* positions_out, vx_block_id* roots_out,
                      size_t cap) {
    if (!it || !path) return fail(VX_E_INVALID, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(VX_E_INVALID, std::string("cannot open ") + path);
    std::vector<u8> raw;
    u8 buf[1 << 16];
    for (size_t got; (got = fread(buf, 1, sizeof(buf), f)) > 0;) raw.insert(raw.end(), buf, buf + got);
    fclose(f);
    VtmCursor r{raw.data(), raw.data() + raw.size()};
    for (int k = 0; k < 12; ++k)
        if (r.byte() != u8("VoxTreeModel"[k])) return fail(VX_E_INVALID, "not a VTM file");  // import.rs:25-27
    const u32 version = (u32(r.byte()) << 8) | r.byte();
    if (version != 0x0100) return fail(VX_E_INVALID, "unsupported VTM version");              // :29-30
    const u32 flags = (u32(r.byte()) << 8) | r.byte();
    if (flags & ~1u) return fail(VX_E_INVALID, "unknown VTM flags");                          // :33 Flags::from_bits
    vx_vtm_info local{};
    local.flags = uint16_t(flags);
    local.max_depth = r.byte();
    const u32 fbits = r.be();
    memcpy(&local.chunk_world_size, &fbits, 4);
    r.be(), r.be();  // reserved
    for (int k = 0; k < 3; ++k) local.world_bounds[k] = int32_t(r.be());
    const u8 name_len = r.byte();
    for (u32 k = 0; k < name_len; ++k) local.name[k] = char(r.byte());
    local.name[name_len] = 0;
    u8 digest[16], check[16];
    for (int k = 0; k < 16; ++k) digest[k] = r.byte();
    const u32 size = r.be();
    if (!r.ok || size_t(r.end - r.p) < size) return fail(VX_E_INVALID, "truncated VTM file");
    std::vector<u8> plain;
    const u8* payload = r.p;
    size_t payload_len = size;
    if (flags & 1u) {
        int rc = zstd_decompress(r.p, size, plain);
        if (rc != VX_OK) return rc;
        payload = plain.data();
        payload_len = plain.size();
    }
    md5_digest(payload, payload_len, check);
    if (memcmp(digest, check, 16) != 0) return fail(VX_E_INVALID, "VTM payload does not match its MD5");  // :87
    if (info) *info = local;
    return vx_model_deserialize(it, payload, payload_len, positions_out, roots_out, cap);
}

}  // extern "C"

// vx_capi_voxelize.inl — host side of the voxeliser (voxelis-voxelize/src/lib.rs:113-249): build_face_to_chunk_map and the
// launch of voxelize_pairs_kernel.  Textually included by vx_capi.cu (inside its extern "C" block).
// ------------------------------------------------------------------------------- voxeliser
// Voxelizer::build_face_to_chunk_map — voxelis-voxelize/src/lib.rs:113-156 (host work in the reference too).
int64_t vx_voxelize_plan(uint8_t max_depth, double chunk_world_size, const double mesh_min[3], size_t n_vertices,
                         const double* vertices, size_t n_faces, const int32_t* faces, int32_t* positions_out,
                         size_t cap_chunks, uint32_t* pair_chunk_out, uint32_t* pair_face_out, size_t cap_pairs,
                         size_t* n_pairs_out) {
    if (!mesh_min || (n_vertices && !vertices) || (n_faces && !faces)) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(max_depth) || !(chunk_world_size > 0)) return fail(VX_E_INVALID, "bad depth or chunk size");
    const int vpa = 1 << max_depth;
    const double voxel_size = chunk_world_size / double(vpa), inv = 1.0 / voxel_size;
    auto as_i32 = [](double v) -> int64_t {
        return v != v ? 0 : v >= 2147483647.0 ? 2147483647 : v <= -2147483648.0 ? -2147483648LL : int64_t(v);
    };
    std::unordered_map<uint64_t, uint32_t> index;
    std::vector<int32_t> pos;
    std::vector<uint32_t> pc, pf;
    pc.reserve(n_faces + n_faces / 2);
    pf.reserve(n_faces + n_faces / 2);
    uint64_t last_key = ~uint64_t(0);  // neighbouring faces of a mesh mostly fall in the chunk of the previous pair
    uint32_t last_index = 0;
    for (size_t f = 0; f < n_faces; ++f) {
        double mn[3], mx[3];
        for (int a = 0; a < 3; ++a) mn[a] = INFINITY, mx[a] = -INFINITY;
        for (int k = 0; k < 3; ++k) {
            const int32_t vi = faces[3 * f + k];
            if (vi < 1 || size_t(vi) > n_vertices) return fail(VX_E_BOUNDS, "face refers to a vertex that does not exist");
            for (int a = 0; a < 3; ++a) {
                const double v = vertices[3 * size_t(vi - 1) + a] - mesh_min[a];
                mn[a] = v < mn[a] ? v : mn[a];
                mx[a] = v > mx[a] ? v : mx[a];
            }
        }
        int64_t c0[3], c1[3];
        for (int a = 0; a < 3; ++a) {
            c0[a] = as_i32(std::floor(mn[a] * inv)) / vpa;  // IVec3 / i32: truncating
            c1[a] = as_i32(std::ceil(mx[a] * inv)) / vpa;
            if (c0[a] < -(1 << 20) || c1[a] >= (1 << 20)) return fail(VX_E_BOUNDS, "mesh spans more than 2^20 chunks");
        }
        for (int64_t cy = c0[1]; cy <= c1[1]; ++cy)
            for (int64_t cz = c0[2]; cz <= c1[2]; ++cz)
                for (int64_t cx = c0[0]; cx <= c1[0]; ++cx) {
                    const uint64_t key = uint64_t(cx + (1 << 20)) | (uint64_t(cy + (1 << 20)) << 21) | (uint64_t(cz + (1 << 20)) << 42);
                    if (key != last_key) {
                        auto ins = index.emplace(key, uint32_t(pos.size() / 3));
                        if (ins.second) {
                            pos.push_back(int32_t(cx));
                            pos.push_back(int32_t(cy));
                            pos.push_back(int32_t(cz));
                        }
                        last_key = key;
                        last_index = ins.first->second;
                    }
                    pc.push_back(last_index);
                    pf.push_back(uint32_t(f));
                }
    }
    if (n_pairs_out) *n_pairs_out = pc.size();
    const size_t n = pos.size() / 3;
    if (positions_out && pair_chunk_out && pair_face_out && cap_chunks >= n && cap_pairs >= pc.size()) {
        if (n) memcpy(positions_out, pos.data(), n * 12);
        if (!pc.empty()) {
            memcpy(pair_chunk_out, pc.data(), pc.size() * 4);
            memcpy(pair_face_out, pf.data(), pf.size() * 4);
        }
    }
    return int64_t(n);
}

// Voxelizer::voxelize_chunk for every planned chunk — voxelis-voxelize/src/lib.rs:159-249.
int vx_voxelize_chunks_device(vx_interner* it, uint8_t max_depth, double chunk_world_size, const double mesh_min[3],
                              size_t n_vertices, const double* vertices, size_t n_faces, const int32_t* faces,
                              size_t n_chunks, const int32_t* positions, size_t n_pairs, const uint32_t* pair_chunk,
                              const uint32_t* pair_face, uint8_t* d_masks, void* d_values, uint8_t* d_has_patches) {
    if (!it || !mesh_min || !d_masks || !d_values) return fail(VX_E_INVALID, "null argument");
    if (!valid_depth(max_depth) || !(chunk_world_size > 0)) return fail(VX_E_INVALID, "bad depth or chunk size");
    if (n_chunks == 0) return VX_OK;
    if (!positions || (n_pairs && (!pair_chunk || !pair_face || !vertices || !faces)))
        return fail(VX_E_INVALID, "null argument");
    if (!is_device_ptr(d_masks) || !is_device_ptr(d_values) || (d_has_patches && !is_device_ptr(d_has_patches)))
        return fail(VX_E_INVALID, "masks, values and has_patches must be device memory");
    if ((reinterpret_cast<uintptr_t>(d_masks) | reinterpret_cast<uintptr_t>(d_values)) & 15)
        return fail(VX_E_INVALID, "device masks/values must be 16-byte aligned");
    if (n_pairs > 0xFFFFFFFFull) return fail(VX_E_INVALID, "more than 2^32 (chunk, face) pairs");
    for (size_t k = 0; k < n_pairs; ++k)
        if (pair_chunk[k] >= n_chunks || pair_face[k] >= n_faces) return fail(VX_E_BOUNDS, "pair out of range");
    for (size_t k = 0; k < n_faces * 3; ++k)
        if (faces[k] < 1 || size_t(faces[k]) > n_vertices) return fail(VX_E_BOUNDS, "face refers to a vertex that does not exist");
    std::lock_guard<std::mutex> lk(it->mu);
    DeviceGuard g(it->device);
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    int rc = ensure_scratch(it, up(n_vertices * 24) + up(n_faces * 12) + up(n_chunks * 12) + 2 * up(n_pairs * 4) + up(n_chunks) + 256, 0);
    if (rc != VX_OK) return rc;
    cudaStream_t s = it->stream;
    u8* p = (u8*)it->scratch;
    auto take = [&](size_t bytes) { u8* r = p; p += up(bytes); return r; };
    double* dv = (double*)take(n_vertices * 24);
    int* df = (int*)take(n_faces * 12);
    int* dp = (int*)take(n_chunks * 12);
    u32* dpc = (u32*)take(n_pairs * 4);
    u32* dpf = (u32*)take(n_pairs * 4);
    u8* dh = d_has_patches ? d_has_patches : take(n_chunks);
    const size_t B = blocks_for_depth(max_depth), esz = dtype_size(it->dtype);
    CU_TRY(cudaMemsetAsync(d_masks, 0, n_chunks * B * 2, s));            // Batch::new, core/batch.rs:63-81
    CU_TRY(cudaMemsetAsync(d_values, 0, n_chunks * B * 8 * esz, s));
    CU_TRY(cudaMemsetAsync(dh, 0, n_chunks, s));
    if (n_pairs) {
        CU_TRY(cudaMemcpyAsync(dv, vertices, n_vertices * 24, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(df, faces, n_faces * 12, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(dp, positions, n_chunks * 12, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(dpc, pair_chunk, n_pairs * 4, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(dpf, pair_face, n_pairs * 4, cudaMemcpyHostToDevice, s));
        const unsigned blocks = unsigned(std::min<size_t>((n_pairs + 7) / 8, size_t(it->sm_count) * 8));
        VoxelizeArgs va{int(max_depth), chunk_world_size, mesh_min[0], mesh_min[1], mesh_min[2], dv, df, dp, dpc, dpf, n_pairs};
        if (it->dtype == VX_U8)
            voxelize_pairs_kernel<u8><<<blocks, 256, 0, s>>>(va, d_masks, (u8*)d_values, dh);
        else
            voxelize_pairs_kernel<int32_t><<<blocks, 256, 0, s>>>(va, d_masks, (int32_t*)d_values, dh);
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaStreamSynchronize(s));
    return VX_OK;
}

// vx_capi_batch.inl — Batch<T> on the host (core/batch.rs:39-45,63-195,211-226): the pinned + mapped batch arena, set / fill /
// clear / assign, the per-unit / per-block occupancy summary and the journal of flagged blocks that vx_apply_batches
// stages from.  Textually included by vx_capi.cu (inside its extern "C" block; uses its fail() / valid_depth() helpers).
// ------------------------------------------------------------------------------- batch
namespace {

// Pinned + mapped host memory for batches, carved into equal slots per (depth, dtype) size class.
// One cudaHostAlloc per batch would cost ~100 us each and thousands of pinned regions; slabs grow
// geometrically to 64 MiB.  Slabs are never returned to the OS (slots are recycled), and are
// deliberately not freed at process exit (the CUDA context may already be gone).
struct BatchArena {
    struct Class {
        std::vector<u8*> free_slots;
        size_t next_slab = 0;
    };
    std::mutex mu;
    std::map<size_t, Class> classes;
    std::map<u8*, std::pair<size_t, u64>> slabs;  // host base -> (bytes, device-visible base)

    u8* take(size_t slot, u64* alias) {
        std::lock_guard<std::mutex> lk(mu);
        Class& c = classes[slot];
        if (c.free_slots.empty()) {
            if (c.next_slab == 0) c.next_slab = std::max<size_t>(slot, size_t(1) << 20);
            size_t count = std::max<size_t>(1, c.next_slab / slot);
            u8* base = nullptr;
            if (cudaHostAlloc((void**)&base, count * slot, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            void* dv = nullptr;
            if (cudaHostGetDevicePointer(&dv, base, 0) != cudaSuccess) {
                cudaGetLastError();
                cudaFreeHost(base);
                return nullptr;
            }
            slabs[base] = {count * slot, u64(reinterpret_cast<uintptr_t>(dv))};
            for (size_t i = count; i-- > 0;) c.free_slots.push_back(base + i * slot);
            c.next_slab = std::min<size_t>(c.next_slab * 2, std::max<size_t>(slot, size_t(64) << 20));
        }
        u8* p = c.free_slots.back();
        c.free_slots.pop_back();
        auto it = slabs.upper_bound(p);
        --it;
        *alias = it->second.second + u64(p - it->first);
        return p;
    }
    void give(u8* p, size_t slot) {
        std::lock_guard<std::mutex> lk(mu);
        classes[slot].free_slots.push_back(p);
    }
};
BatchArena& arena() {
    static BatchArena* a = new BatchArena();
    return *a;
}

inline void batch_touch(vx_batch* b, size_t block) {
    size_t u = block / b->unit_blocks;
    b->touched[u >> 6] |= uint64_t(1) << (u & 63);
    b->occ[block >> 3] |= u8(1u << (block & 7));
}
// the journal's copy of block p follows the dense array (called after every write to a block that is, or now has to
// be, in the journal)
inline void journal_update(vx_batch* b, size_t p, bool enter) {
    if (!b->journal_ok) return;
    uint32_t s = b->jmap[p];
    if (!s) {
        if (!enter) return;
        if (b->jcount == b->jcap) {
            b->journal_ok = false;  // too dense for a journal: the occupancy bitmap serves this batch
            return;
        }
        s = ++b->jcount;
        b->jmap[p] = uint16_t(s);
        b->jblock[s - 1] = uint16_t(p);
    }
    const size_t vsz = 8 * dtype_size(b->dtype);
    memcpy((u8*)b->jvals + size_t(s - 1) * vsz, (const u8*)b->values + p * vsz, vsz);
}
inline void journal_reset(vx_batch* b) {
    if (b->jmap) {
        if (b->journal_ok)
            for (uint32_t k = 0; k < b->jcount; ++k) b->jmap[b->jblock[k]] = 0;
        else
            memset(b->jmap, 0, b->blocks * 2);
    }
    b->jcount = 0;
    b->journal_ok = b->jmap != nullptr && !b->raw_exposed;
}
// Rebuilds the occupancy summary from the arrays (after the caller bulk-wrote them).
void batch_rescan(vx_batch* b) {
    memset(b->touched, 0, sizeof(b->touched));
    memset(b->occ, 0, (b->blocks + 7) / 8);
    for (uint32_t u = 0; u < b->units; ++u) {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(b->masks + size_t(u) * b->unit_blocks * 2);
        uint64_t any = 0;
        for (size_t k = 0, e = size_t(b->unit_blocks) * 2 / 8; k < e; ++k) any |= w[k];
        if (!(any & 0x00FF00FF00FF00FFull)) continue;  // set_mask bytes
        b->touched[u >> 6] |= uint64_t(1) << (u & 63);
        for (size_t p = size_t(u) * b->unit_blocks, e = p + b->unit_blocks; p < e; ++p)
            if (b->masks[2 * p]) b->occ[p >> 3] |= u8(1u << (p & 7));
    }
    journal_reset(b);
    if (b->journal_ok)
        for (size_t p = 0; p < b->blocks && b->journal_ok; ++p)
            if (b->occ[p >> 3] >> (p & 7) & 1) journal_update(b, p, true);
}

}  // namespace

vx_batch* vx_batch_create(uint8_t max_depth, vx_dtype dtype) {
    if (!valid_depth(max_depth) || (dtype != VX_U8 && dtype != VX_I32)) {
        fail(VX_E_INVALID, "vx_batch_create: max_depth must be in [2,7] and dtype u8/i32");
        return nullptr;
    }
    vx_batch* b = new vx_batch();
    b->depth = max_depth;
    b->dtype = dtype;
    b->blocks = blocks_for_depth(max_depth);
    b->has_fill = false;
    b->fill = 0;
    b->has_patches = false;
    b->raw_exposed = false;
    b->unit_blocks = uint32_t(std::min<size_t>(b->blocks, 512));
    b->units = uint32_t(b->blocks / b->unit_blocks);
    memset(b->touched, 0, sizeof(b->touched));
    // pinned + mapped so apply can read the batch in place over PCIe; no device, no batch
    size_t mb = b->blocks * 2, vb = b->blocks * 8 * dtype_size(dtype);
    const size_t ob = ((b->blocks + 7) / 8 + 3 + 15) & ~size_t(15);  // +3: the staging kernel reads the bitmap as 32-bit words
    // journal (see vx_batch): a quarter of the blocks at most, D <= 6 (block indices are 16 bits)
    b->jcap = max_depth <= 6 && b->blocks >= 64 ? uint32_t(b->blocks / 4) : 0;
    const size_t jv = size_t(b->jcap) * 8 * dtype_size(dtype), jb = (size_t(b->jcap) * 2 + 15) & ~size_t(15),
                 jm = b->jcap ? b->blocks * 2 : 0;
    b->slot_bytes = (mb + vb + ob + jv + jb + jm + 255) & ~size_t(255);
    b->masks = arena().take(b->slot_bytes, &b->alias);
    if (!b->masks) {
        fail(VX_E_CUDA, "vx_batch_create: cudaHostAlloc failed (no CUDA device?)");
        delete b;
        return nullptr;
    }
    b->values = b->masks + mb;
    b->occ = b->masks + mb + vb;
    b->jvals = b->jcap ? b->masks + mb + vb + ob : nullptr;
    b->jblock = b->jcap ? reinterpret_cast<uint16_t*>(b->masks + mb + vb + ob + jv) : nullptr;
    b->jmap = b->jcap ? reinterpret_cast<uint16_t*>(b->masks + mb + vb + ob + jv + jb) : nullptr;
    b->jcount = 0;
    b->journal_ok = b->jcap != 0;
    memset(b->masks, 0, b->slot_bytes);
    return b;
}
void vx_batch_destroy(vx_batch* b) {
    if (!b) return;
    arena().give(b->masks, b->slot_bytes);
    delete b;
}

static inline u32 spread10(u32 v) {  // utils/common.rs:24-55
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x30000FF;
    v = (v | (v << 8)) & 0x300F00F;
    v = (v | (v << 4)) & 0x30C30C3;
    v = (v | (v << 2)) & 0x9249249;
    return v;
}

int vx_batch_set(vx_batch* b, int x, int y, int z, int64_t voxel) {
    if (!b) return fail(VX_E_INVALID, "null batch");
    int n = 1 << b->depth;
    if (x < 0 || y < 0 || z < 0 || x >= n || y >= n || z >= n) return fail(VX_E_BOUNDS, "position out of bounds");
    u32 full = spread10(u32(x)) | (spread10(u32(y)) << 1) | (spread10(u32(z)) << 2);
    size_t p = full >> 3;
    u32 i = full & 7;
    u8 bit = u8(1u << i);
    bool nonzero;
    if (b->dtype == VX_U8) {
        u8 v = u8(voxel);
        nonzero = v != 0;
        ((u8*)b->values)[p * 8 + i] = v;
    } else {
        int32_t v = int32_t(voxel);
        nonzero = v != 0;
        ((int32_t*)b->values)[p * 8 + i] = v;
    }
    if (nonzero) {  // batch.rs:162-168
        b->masks[2 * p] |= bit;
        b->masks[2 * p + 1] &= u8(~bit);
        batch_touch(b, p);
    } else {
        b->masks[2 * p] &= u8(~bit);
        b->masks[2 * p + 1] |= bit;
    }
    journal_update(b, p, nonzero);
    b->has_patches = true;
    return 1;
}
int vx_batch_clear(vx_batch* b) {
    if (!b) return fail(VX_E_INVALID, "null batch");
    const size_t vsz = 8 * dtype_size(b->dtype);
    if (b->raw_exposed) {
        memset(b->masks, 0, b->blocks * 2 + b->blocks * vsz);
    } else if (b->has_patches) {
        // only set() wrote into the arrays.  Units with a set bit are known; a clear_mask bit (value 0,
        // batch.rs:165-168) can sit anywhere, so the masks are wiped whole and the values per touched unit.
        memset(b->masks, 0, b->blocks * 2);
        for (uint32_t u = 0; u < b->units; ++u)
            if (b->touched[u >> 6] >> (u & 63) & 1)
                memset((u8*)b->values + size_t(u) * b->unit_blocks * vsz, 0, size_t(b->unit_blocks) * vsz);
    }
    memset(b->touched, 0, sizeof(b->touched));
    memset(b->occ, 0, (b->blocks + 7) / 8);
    journal_reset(b);
    b->has_fill = false;
    b->fill = 0;
    b->has_patches = false;
    return VX_OK;
}
int vx_batch_fill(vx_batch* b, int64_t value) {
    int rc = vx_batch_clear(b);
    if (rc != VX_OK) return rc;
    b->has_fill = true;
    b->fill = b->dtype == VX_U8 ? int64_t(u8(value)) : int64_t(int32_t(value));
    return VX_OK;
}
uint8_t* vx_batch_masks(vx_batch* b) {
    if (b) b->raw_exposed = true, b->journal_ok = false;  // the caller may write behind the journal's back
    return b ? b->masks : nullptr;
}
void* vx_batch_values(vx_batch* b) {
    if (b) b->raw_exposed = true, b->journal_ok = false;
    return b ? b->values : nullptr;
}
size_t vx_batch_blocks(const vx_batch* b) { return b ? b->blocks : 0; }
int vx_batch_to_fill(const vx_batch* b, int64_t* out) {
    if (!b) return fail(VX_E_INVALID, "null batch");
    if (b->has_fill && out) *out = b->fill;
    return b->has_fill ? 1 : 0;
}
size_t vx_batch_size(const vx_batch* b) {  // batch.rs:110-121
    if (!b) return 0;
    size_t n = 0;
    for (size_t p = 0; p < b->blocks; ++p) n += (b->masks[2 * p] | b->masks[2 * p + 1]) != 0;
    return n;
}
int vx_batch_has_patches(const vx_batch* b) { return b && b->has_patches; }
void vx_batch_mark_patched(vx_batch* b) {
    if (!b) return;
    b->has_patches = true;
    batch_rescan(b);
}
int vx_batch_set_many(vx_batch* b, size_t n, const int32_t* xyz, const int64_t* voxels) {
    if (!b || (n && (!xyz || !voxels))) return fail(VX_E_INVALID, "null argument");
    for (size_t i = 0; i < n; ++i) {
        int rc = vx_batch_set(b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], voxels[i]);
        if (rc < 0) return rc;
    }
    return int(n > 0);
}
int vx_batch_assign(vx_batch* b, const uint8_t* masks, const void* values) {
    if (!b || !masks || !values) return fail(VX_E_INVALID, "null argument");
    // same state as replaying Batch::set for every recorded voxel (batch.rs:145-175): a value is kept only
    // under its set bit, and a set bit whose value is the default cannot come from set() and is dropped
    const size_t B = b->blocks;
    bool any = false;
    for (size_t p = 0; p < B; ++p) {
        u8 set = masks[2 * p], keep = 0;
        if (b->dtype == VX_U8) {
            const u8* v = (const u8*)values + p * 8;
            u8* o = (u8*)b->values + p * 8;
            for (int i = 0; i < 8; ++i) {
                o[i] = (set >> i & 1) ? v[i] : 0;
                keep |= u8((o[i] != 0) << i);
            }
        } else {
            const int32_t* v = (const int32_t*)values + p * 8;
            int32_t* o = (int32_t*)b->values + p * 8;
            for (int i = 0; i < 8; ++i) {
                o[i] = (set >> i & 1) ? v[i] : 0;
                keep |= u8((o[i] != 0) << i);
            }
        }
        b->masks[2 * p] = keep;
        b->masks[2 * p + 1] = masks[2 * p + 1];
        any = any || (masks[2 * p] | masks[2 * p + 1]) != 0;
    }
    b->has_fill = false;
    b->fill = 0;
    b->has_patches = any;
    batch_rescan(b);
    return VX_OK;
}
int vx_batch_touched_units(const vx_batch* b) {
    if (!b) return fail(VX_E_INVALID, "null batch");
    int c = 0;
    for (uint64_t w : b->touched) c += __builtin_popcountll(w);
    return c;
}
uint8_t vx_batch_max_depth(const vx_batch* b) { return b ? b->depth : 0; }
vx_dtype vx_batch_dtype(const vx_batch* b) { return b ? b->dtype : VX_U8; }

// voxelis_b200.hpp — header-only C++ mirror of the reference's Rust API over the C ABI.
//
// The reference is compiled Rust; its toolchain is not available in this image, so the host side
// above the C ABI is C++.  Names, argument meaning and error behaviour follow the Rust traits
// (voxelis/src/spatial/voxops.rs:8-35) so call sites read like the reference's:
//
//     auto interner = voxelis::VoxInterner<uint8_t>::with_memory_budget(256u << 20);   // README.md:59-60
//     voxelis::VoxTree<uint8_t> tree(5);                                               // MaxDepth::new(5)
//     auto batch = tree.create_batch();                                                // voxtree.rs:296
//     batch.set(interner, {1, 2, 3}, 7);                                               // batch.rs:211
//     tree.apply_batch(interner, batch);                                               // voxtree.rs:303
//     std::optional<uint8_t> v = tree.get(interner, {1, 2, 3});                        // voxtree.rs:145
//
// Reference panics (assert!/panic!) surface as voxelis::Error carrying the vx_status code.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "voxelis_b200.h"

namespace voxelis {

struct IVec3 {
    int x, y, z;
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace detail {
template <class T>
constexpr vx_dtype dtype_of() {
    static_assert(std::is_same<T, uint8_t>::value || std::is_same<T, int32_t>::value,
                  "VoxelTrait impls available on this path: u8 and i32 (core/voxel.rs:77,93)");
    return std::is_same<T, uint8_t>::value ? VX_U8 : VX_I32;
}
inline int check(int rc) {
    if (rc < 0) throw Error(rc, vx_last_error());
    return rc;
}
}  // namespace detail

/// BlockId accessors (core/block_id.rs:240-401)
struct BlockId {
    vx_block_id raw;
    static constexpr vx_block_id EMPTY = VX_BLOCK_EMPTY;
    static constexpr vx_block_id INVALID = VX_BLOCK_INVALID;
    uint32_t index() const { return uint32_t(raw); }
    uint16_t generation() const { return uint16_t((raw >> 32) & 0x7FFF); }
    bool is_leaf() const { return (raw >> 63) == 1; }
    bool is_branch() const { return (raw >> 63) == 0; }
    bool is_empty() const { return raw == 0; }
    uint8_t types() const { return uint8_t(raw >> 55); }
    uint8_t mask() const { return uint8_t(raw >> 47); }
};

template <class T>
class VoxTree;
template <class T>
class Batch;

/// VoxInterner<T> — interner/mod.rs:25-40
template <class T>
class VoxInterner {
public:
    static VoxInterner with_memory_budget(size_t bytes, int device = 0) {  // interner/mod.rs:45
        vx_interner* h = vx_interner_create(bytes, detail::dtype_of<T>(), device);
        if (!h) throw Error(VX_E_BUDGET, vx_last_error());
        return VoxInterner(h);
    }
    VoxInterner(VoxInterner&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    VoxInterner(const VoxInterner&) = delete;
    ~VoxInterner() { vx_interner_destroy(h_); }
    uint32_t get_ref(BlockId id) const {  // mod.rs:218
        uint32_t r = 0;
        detail::check(vx_interner_get_ref(h_, id.raw, &r));
        return r;
    }
    T get_value(BlockId id) const {  // mod.rs:166
        int64_t v = 0;
        detail::check(vx_interner_get_value(h_, id.raw, &v));
        return T(v);
    }
    vx_stats stats() const {  // interner/stats.rs
        vx_stats s{};
        detail::check(vx_interner_stats(h_, &s));
        return s;
    }
    size_t capacity() const { return vx_interner_capacity(h_); }
    void reset() { detail::check(vx_interner_reset(h_)); }
    vx_interner* raw() const { return h_; }

private:
    explicit VoxInterner(vx_interner* h) : h_(h) {}
    vx_interner* h_;
};

/// Batch<T> — core/batch.rs:39-45
template <class T>
class Batch {
public:
    explicit Batch(uint8_t max_depth) : h_(vx_batch_create(max_depth, detail::dtype_of<T>())) {  // batch.rs:63
        if (!h_) throw Error(VX_E_INVALID, vx_last_error());
    }
    Batch(Batch&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    Batch(const Batch&) = delete;
    ~Batch() { vx_batch_destroy(h_); }
    bool set(VoxInterner<T>&, IVec3 p, T voxel) { return detail::check(vx_batch_set(h_, p.x, p.y, p.z, int64_t(voxel))) == 1; }
    bool just_set(IVec3 p, T voxel) { return detail::check(vx_batch_set(h_, p.x, p.y, p.z, int64_t(voxel))) == 1; }
    void fill(VoxInterner<T>&, T value) { detail::check(vx_batch_fill(h_, int64_t(value))); }  // batch.rs:218
    void clear(VoxInterner<T>&) { detail::check(vx_batch_clear(h_)); }                          // batch.rs:223
    uint8_t* masks() { return vx_batch_masks(h_); }                                             // batch.rs:86 ([B][2])
    T* values() { return static_cast<T*>(vx_batch_values(h_)); }                                // batch.rs:97 ([B][8])
    std::optional<T> to_fill() const {                                                          // batch.rs:108
        int64_t v = 0;
        return detail::check(vx_batch_to_fill(h_, &v)) == 1 ? std::optional<T>(T(v)) : std::nullopt;
    }
    size_t size() const { return vx_batch_size(h_); }
    bool has_patches() const { return vx_batch_has_patches(h_) != 0; }
    void mark_patched() { vx_batch_mark_patched(h_); }  // after bulk-writing masks()/values()
    vx_batch* raw() const { return h_; }

private:
    vx_batch* h_;
};

/// VoxTree<T> — spatial/voxtree.rs:108-113 with VoxOpsRead / VoxOpsBulkWrite / VoxOpsBatch / state traits
template <class T>
class VoxTree {
public:
    explicit VoxTree(uint8_t max_depth) : h_(vx_tree_create(max_depth)) {  // voxtree.rs:116
        if (!h_) throw Error(VX_E_INVALID, vx_last_error());
    }
    VoxTree(VoxTree&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    VoxTree(const VoxTree&) = delete;
    ~VoxTree() { vx_tree_destroy(h_); }
    Batch<T> create_batch() const { return Batch<T>(vx_tree_max_depth(h_)); }                  // voxtree.rs:296
    bool apply_batch(VoxInterner<T>& i, const Batch<T>& b) {                                    // voxtree.rs:303
        return detail::check(vx_tree_apply_batch(i.raw(), h_, b.raw())) == 1;
    }
    std::optional<T> get(const VoxInterner<T>& i, IVec3 p) const {                              // voxtree.rs:145
        int64_t v = 0;
        return detail::check(vx_tree_get(i.raw(), h_, p.x, p.y, p.z, &v)) == 1 ? std::optional<T>(T(v)) : std::nullopt;
    }
    std::vector<T> to_vec(const VoxInterner<T>& i) const {                                      // utils/common.rs:158
        size_t n = size_t(1) << vx_tree_max_depth(h_);
        std::vector<T> dense(n * n * n);
        detail::check(vx_tree_to_vec(i.raw(), h_, dense.data()));
        return dense;
    }
    /// to_vec(interner, root, max_depth.for_lod(lod)) — world/voxchunk.rs:267, core/max_depth.rs:137-140
    std::vector<T> to_vec(const VoxInterner<T>& i, uint8_t lod) const {
        const uint8_t d = vx_tree_max_depth(h_);
        size_t n = size_t(1) << (d > lod ? d - lod : 0);
        std::vector<T> dense(n * n * n);
        vx_block_id root = vx_tree_root_id(h_);
        detail::check(vx_roots_to_vec_lod(i.raw(), d, lod, 1, &root, dense.data()));
        return dense;
    }
    void fill(VoxInterner<T>& i, T value) { detail::check(vx_tree_fill(i.raw(), h_, int64_t(value))); }  // voxtree.rs:264
    void clear(VoxInterner<T>& i) { detail::check(vx_tree_clear(i.raw(), h_)); }                         // voxtree.rs:283
    BlockId get_root_id() const { return BlockId{vx_tree_root_id(h_)}; }
    bool is_empty() const { return vx_tree_is_empty(h_) != 0; }
    bool is_leaf() const { return vx_tree_is_leaf(h_) != 0; }
    bool is_dirty() const { return vx_tree_is_dirty(h_) != 0; }
    void mark_dirty() { vx_tree_mark_dirty(h_); }
    void clear_dirty() { vx_tree_clear_dirty(h_); }
    uint8_t max_depth() const { return vx_tree_max_depth(h_); }
    uint32_t voxels_per_axis() const { return vx_tree_voxels_per_axis(h_); }
    vx_tree* raw() const { return h_; }

private:
    vx_tree* h_;
};

/// New multi-chunk entry: replaces the serial loop of voxelis-voxelize/src/lib.rs:357-361.
template <class T>
std::vector<bool> apply_batches(VoxInterner<T>& interner, const std::vector<VoxTree<T>*>& trees,
                                const std::vector<const Batch<T>*>& batches) {
    std::vector<vx_tree*> t;
    std::vector<const vx_batch*> b;
    for (auto* x : trees) t.push_back(x->raw());
    for (auto* x : batches) b.push_back(x->raw());
    std::vector<uint8_t> changed(t.size());
    detail::check(vx_apply_batches(interner.raw(), t.data(), b.data(), t.size(), changed.data()));
    return std::vector<bool>(changed.begin(), changed.end());
}

}  // namespace voxelis

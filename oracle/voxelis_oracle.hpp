// voxelis_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A C++17 restatement of the batched SVO-DAG build path of WildPixelGames/voxelis
// v25.4.0 (Rust).  It exists only to check the CUDA path in voxelis_b200/ and to be
// timed as the CPU baseline by bench.py.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product path never
// calls into this file and fails loudly when the CUDA library is missing.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/voxelis/src unless stated).
//
// Parity status of the oracle itself:
//   * The reference is Rust-only; this image has no cargo/rustc, so the reference cannot
//     be run here and oracle/_ref does not exist ("reference unbuildable here").
//   * The oracle is PINNED against every known-answer the reference's own test-suite
//     holds for this path (tests/test_oracle_reference_tests.py ports
//     spatial/voxtree.rs:1345-2185, core/block_id.rs:411,474-609,
//     core/max_depth.rs:167-172) and cross-checked against an independent canonical
//     builder (tests/canonical.py).
//   * Hash VALUES are "parity unpinned": the reference hashes with rustc-hash 2.1.1
//     FxHasher (Cargo.lock; call sites interner/hash.rs:42,57,70), which is not vendored
//     under /root/reference and is pinned by no reference test.  Hashes never escape the
//     interner and ids are compared up to permutation, so this does not affect results.
//     fx_* below restates the published FxHasher algorithm for the stored `hashes` pool.
//   * Deliberate deviation: the reference keys its pattern maps on the 64-bit hash only
//     (interner/hash.rs:15-38; equality is only debug_assert'ed, interner/mod.rs:749-765).
//     The oracle compares full keys, so it differs from the reference only where the
//     reference would silently alias two different nodes on a 64-bit hash collision.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace vxo {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

// ---------------------------------------------------------------------------------
// BlockId — core/block_id.rs:18-30 (shifts), :94-130 (consts), :216-226 (packing),
// :240-401 (accessors).
// ---------------------------------------------------------------------------------
constexpr u64 ID_INVALID = ~u64(0);      // block_id.rs:96
constexpr u64 ID_EMPTY = 0;              // block_id.rs:112
constexpr u16 MAX_GENERATION = 0x7FFE;   // block_id.rs:130
constexpr int MAX_CHILDREN = 8;          // interner/consts.rs:6

struct RefPanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline u64 id_pack(u32 index, u16 gen, u8 types, u8 mask, bool leaf) {
    if (gen > MAX_GENERATION) throw RefPanic("Generation overflow");  // :217
    return (u64(leaf) << 63) | (u64(types) << 55) | (u64(mask) << 47) |
           ((u64(gen) & 0x7FFF) << 32) | u64(index);
}
inline u64 id_leaf(u32 index, u16 gen) { return id_pack(index, gen, 0, 0, true); }
inline u64 id_branch(u32 index, u16 gen, u8 t, u8 m) { return id_pack(index, gen, t, m, false); }
inline u32 id_index(u64 id) { return u32(id & 0xFFFFFFFFull); }
inline u16 id_gen(u64 id) { return u16((id >> 32) & 0x7FFF); }
inline bool id_is_leaf(u64 id) { return (id >> 63) == 1; }
inline bool id_is_branch(u64 id) { return (id >> 63) == 0; }
inline bool id_is_empty(u64 id) { return id == 0; }
inline u8 id_types(u64 id) {
    if (!id_is_branch(id)) throw RefPanic("Cannot get types from a leaf node");  // :273-276
    return u8((id >> 55) & 0xFF);
}
inline u8 id_mask(u64 id) {
    if (!id_is_branch(id)) throw RefPanic("Cannot get mask from a leaf node");  // :296-299
    return u8((id >> 47) & 0xFF);
}

// ---------------------------------------------------------------------------------
// Morton path — utils/common.rs:24-55.  x -> bit 3n, y -> 3n+1, z -> 3n+2, 10 bits/axis.
// ---------------------------------------------------------------------------------
inline u32 spread10(u32 v) {
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x30000FF;
    v = (v | (v << 8)) & 0x300F00F;
    v = (v | (v << 4)) & 0x30C30C3;
    v = (v | (v << 2)) & 0x9249249;
    return v;
}
inline u32 encode_child_index_path(int x, int y, int z) {
    return spread10(u32(x)) | (spread10(u32(y)) << 1) | (spread10(u32(z)) << 2);
}

// PATH_MASKS[max_depth][level] — spatial/voxtree.rs:41-105: the top (level+1) 3-bit
// groups of a 3*max_depth-bit path.  The table stops at max_depth 6; the formula below
// reproduces every tabulated row and extends it to max_depth 7 (extension).
inline u32 path_mask(int max_depth, int level) {
    if (level >= max_depth) return 0;
    u32 ones = (u32(1) << (3 * (level + 1))) - 1;
    return ones << (3 * (max_depth - level - 1));
}

// ---------------------------------------------------------------------------------
// FxHasher (rustc-hash 2.1.1, published algorithm; parity unpinned, see header).
// ---------------------------------------------------------------------------------
struct Fx {
    u64 h = 0;
    static constexpr u64 K = 0xf1357aea2e62a9c5ull;
    void add(u64 v) { h = (h + v) * K; }
    u64 finish() const { return (h << 26) | (h >> 38); }
};
template <class T>
inline u64 fx_leaf_hash(T v) {  // interner/hash.rs:54-64
    Fx f;
    f.add(0);  // NODE_TYPE_LEAF, consts.rs:7
    f.add(u64(typename std::make_unsigned<T>::type(v)));
    return f.finish();
}
inline u64 fx_branch_hash(const u64* c, u8 types, u8 mask) {  // interner/hash.rs:67-81
    Fx f;
    f.add(1);  // NODE_TYPE_BRANCH
    f.add((u64(types) << 8) | mask);
    for (int i = 0; i < 8; ++i) f.add(c[i]);
    return f.finish();
}
inline u64 fx_empty_branch_hash() {  // interner/hash.rs:41-51
    Fx f;
    f.add(1);
    for (int i = 0; i < 8; ++i) f.add(0);
    return f.finish();
}

// ---------------------------------------------------------------------------------
// calc_average — core/voxel.rs:96-141 (mode of the 8 child values; ties: a non-default
// value beats the default, otherwise first occurrence wins).
// ---------------------------------------------------------------------------------
template <class T>
inline T calc_average(const T* c) {
    T values[8];
    size_t counts[8];
    int unique = 0;
    for (int k = 0; k < 8; ++k) {
        int i = 0;
        while (i < unique) {
            if (values[i] == c[k]) {
                counts[i] += 1;
                break;
            }
            ++i;
        }
        if (i == unique) {
            values[unique] = c[k];
            counts[unique] = 1;
            ++unique;
        }
    }
    if (unique == 0) return T(0);
    int max_i = 0;
    size_t max_cnt = counts[0];
    for (int i = 1; i < unique; ++i) {
        bool def_i = values[i] == T(0), def_max = values[max_i] == T(0);
        if (counts[i] > max_cnt || (counts[i] == max_cnt && def_max && !def_i)) {
            max_cnt = counts[i];
            max_i = i;
        }
    }
    return values[max_i];
}

// ---------------------------------------------------------------------------------
// InternerStats — interner/stats.rs:1-28 (feature memory_stats); update sites cited
// where they are bumped.
// ---------------------------------------------------------------------------------
struct Stats {
    u64 requested_budget = 0, actual_budget = 0, node_size = 0, nodes_capacity = 0;
    u64 total_allocations = 0, total_deallocations = 0, allocated_nodes = 0, recycled_nodes = 0;
    u64 alive_nodes = 0, patterns = 0;
    u64 total_cache_hits = 0, total_cache_misses = 0;
    u64 branch_cache_hits = 0, branch_cache_misses = 0, leaf_cache_hits = 0, leaf_cache_misses = 0;
    u64 collapsed_branches = 0, leaf_nodes = 0, branch_nodes = 0;
    u64 max_alive_nodes = 0, max_node_id = 0, max_branch_ref_count = 0, max_leaf_ref_count = 0;
    u64 max_generation = 0, generations_overflows = 0;
};

// Open-addressing hash -> id map standing in for the reference's
// HashMap<u64, BlockId, IdentityHasher> (interner/hash.rs:15-38).  Linear probing with
// back-shift deletion; equality is decided by the caller's full-key predicate.
struct PatternMap {
    struct Slot {
        u64 hash;
        u64 id;  // ID_INVALID == vacant
    };
    std::vector<Slot> slots;
    size_t mask = 0, count = 0;
    explicit PatternMap(size_t cap = 16384) {  // INITIAL_CAPACITY, interner/mod.rs:43
        size_t n = 1;
        while (n < cap * 2) n <<= 1;
        slots.assign(n, Slot{0, ID_INVALID});
        mask = n - 1;
    }
    static size_t mix(u64 h) {
        h ^= h >> 32;
        h *= 0x9E3779B97F4A7C15ull;
        return size_t(h ^ (h >> 29));
    }
    template <class Eq>
    Slot* find(u64 hash, Eq eq) {
        size_t i = mix(hash) & mask;
        while (slots[i].id != ID_INVALID) {
            if (slots[i].hash == hash && eq(slots[i].id)) return &slots[i];
            i = (i + 1) & mask;
        }
        return nullptr;
    }
    void grow() {
        std::vector<Slot> old;
        old.swap(slots);
        slots.assign(old.size() * 2, Slot{0, ID_INVALID});
        mask = slots.size() - 1;
        for (auto& s : old)
            if (s.id != ID_INVALID) {
                size_t i = mix(s.hash) & mask;
                while (slots[i].id != ID_INVALID) i = (i + 1) & mask;
                slots[i] = s;
            }
    }
    void insert(u64 hash, u64 id) {
        if ((count + 1) * 2 > slots.size()) grow();
        size_t i = mix(hash) & mask;
        while (slots[i].id != ID_INVALID) i = (i + 1) & mask;
        slots[i] = Slot{hash, id};
        ++count;
    }
    void erase_id(u64 hash, u64 id) {
        size_t i = mix(hash) & mask;
        while (slots[i].id != ID_INVALID) {
            if (slots[i].hash == hash && slots[i].id == id) break;
            i = (i + 1) & mask;
        }
        if (slots[i].id == ID_INVALID) return;
        // back-shift deletion
        size_t j = i;
        for (;;) {
            j = (j + 1) & mask;
            if (slots[j].id == ID_INVALID) break;
            size_t home = mix(slots[j].hash) & mask;
            bool between = (i <= j) ? (home > i && home <= j) : (home > i || home <= j);
            if (!between) {
                slots[i] = slots[j];
                i = j;
            }
        }
        slots[i].id = ID_INVALID;
        --count;
    }
};

// ---------------------------------------------------------------------------------
// VoxInterner<T> — interner/mod.rs:25-40 (struct), :45-155 (with_memory_budget),
// :158-164 (node_size), pools = voxelis-memory/src/pool_allocator_lite.rs:29-81
// (one zeroed slab per field -> calloc here, lazily paged like alloc_zeroed).
// ---------------------------------------------------------------------------------
template <class T>
struct Interner {
    PatternMap patterns[2];  // [0]=branch, [1]=leaf — consts.rs:10-11
    std::vector<u32> free_indices;
    u32 next_index = 1;
    u32* ref_counts = nullptr;
    u16* generations = nullptr;
    u64 (*children)[8] = nullptr;
    T* values = nullptr;
    u64* hashes = nullptr;
    size_t capacity = 0;
    u64 empty_branch_hash = 0;
    Stats stats;

    static constexpr size_t node_size() { return 4 + 2 + 64 + sizeof(T) + 8; }  // mod.rs:158-164

    explicit Interner(size_t requested_budget) {  // mod.rs:45-155
        size_t cap = requested_budget / node_size();
        size_t actual = cap * node_size();
        if (cap == 0 || actual == 0) throw RefPanic("Requested budget is too small");  // :63-64
        if (cap > 0xFFFFFFFFull) throw RefPanic("Requested budget is too large");       // :65-68
        capacity = cap;
        ref_counts = (u32*)calloc(cap, sizeof(u32));
        generations = (u16*)calloc(cap, sizeof(u16));
        children = (u64(*)[8])calloc(cap, 64);
        values = (T*)calloc(cap, sizeof(T));
        hashes = (u64*)calloc(cap, sizeof(u64));
        if (!ref_counts || !generations || !children || !values || !hashes)
            throw std::bad_alloc();
        empty_branch_hash = fx_empty_branch_hash();
        hashes[0] = empty_branch_hash;                 // :92-96 slot 0 = permanent empty branch
        patterns[0].insert(empty_branch_hash, ID_EMPTY);  // :97
        next_index = 1;                                // :101
        stats.requested_budget = requested_budget;     // :111-137
        stats.actual_budget = actual;
        stats.node_size = node_size();
        stats.nodes_capacity = cap;
        stats.total_allocations = 1;
        stats.allocated_nodes = 1;
        stats.alive_nodes = 1;
        stats.patterns = 1;
        stats.branch_nodes = 1;
    }
    ~Interner() {
        free(ref_counts);
        free(generations);
        free(children);
        free(values);
        free(hashes);
    }
    Interner(const Interner&) = delete;
    Interner& operator=(const Interner&) = delete;

    // get_next_index_macro! — interner/macros.rs:1-41
    u32 next_free_index() {
        if (!free_indices.empty()) {
            u32 i = free_indices.back();
            free_indices.pop_back();
            stats.alive_nodes += 1;
            if (stats.alive_nodes > stats.max_alive_nodes) stats.max_alive_nodes = stats.alive_nodes;
            stats.recycled_nodes -= 1;
            stats.total_allocations += 1;
            return i;
        } else if (next_index < u32(capacity)) {
            u32 i = next_index++;
            stats.alive_nodes += 1;
            if (stats.alive_nodes > stats.max_alive_nodes) stats.max_alive_nodes = stats.alive_nodes;
            stats.allocated_nodes += 1;
            stats.total_allocations += 1;
            if (i > stats.max_node_id) stats.max_node_id = i;
            return i;
        }
        throw RefPanic("Out of memory");  // macros.rs:38
    }

    bool is_valid(u64 id) const { return generations[id_index(id)] == id_gen(id); }  // mod.rs:1004-1008
    T get_value(u64 id) const { return values[id_index(id)]; }                       // :166-176
    u64 get_child_id(u64 id, int i) const { return children[id_index(id)][i]; }      // :205-216
    u32 get_ref(u64 id) const { return ref_counts[id_index(id)]; }                   // :218-228

    void note_ref(u64 id) {
        u64 r = ref_counts[id_index(id)];
        if (id_is_branch(id)) {
            if (r > stats.max_branch_ref_count) stats.max_branch_ref_count = r;
        } else if (r > stats.max_leaf_ref_count)
            stats.max_leaf_ref_count = r;
    }
    void inc_ref(u64 id) {  // mod.rs:230-256
        ref_counts[id_index(id)] += 1;
        note_ref(id);
    }
    void inc_ref_by(u64 id, u32 n) {  // mod.rs:342-370
        ref_counts[id_index(id)] += n;
        note_ref(id);
    }
    void remove_pattern(u64 id) {
        patterns[id_is_leaf(id) ? 1 : 0].erase_id(hashes[id_index(id)], id);
        stats.patterns -= 1;
    }
    bool dec_ref(u64 id) {  // mod.rs:258-289
        u32 i = id_index(id);
        ref_counts[i] -= 1;
        if (ref_counts[i] == 0) {
            remove_pattern(id);
            recycle(id);
            return true;
        }
        return false;
    }
    void dec_ref_by(u64 id, u32 n) {  // mod.rs:372-417
        u32 i = id_index(id);
        ref_counts[i] -= n;
        if (ref_counts[i] == 0) {
            remove_pattern(id);
            recycle(id);
        }
    }
    void dec_child_refs(const u64* c) {  // mod.rs:536-564
        for (int i = 0; i < 8; ++i)
            if (!id_is_empty(c[i])) dec_ref(c[i]);
    }
    // dec_ref_recursive — mod.rs:419-534 (FIFO work list; the reference preallocates
    // 32768 entries, consts.rs:15 — a growable vector here).
    void dec_ref_recursive(u64 root) {
        std::vector<u64> stack;
        stack.push_back(root);
        size_t read = 0;
        while (read < stack.size()) {
            u64 cur = stack[read++];
            u32 ci = id_index(cur);
            ref_counts[ci] -= 1;
            if (ref_counts[ci] == 0) {
                for (int k = 0; k < 8; ++k) {
                    u64 ch = children[ci][k];
                    if (!id_is_empty(ch)) {
                        u32& rc = ref_counts[id_index(ch)];
                        if (rc > 1)
                            rc -= 1;
                        else
                            stack.push_back(ch);
                    }
                }
                remove_pattern(cur);
                recycle(cur);
            }
        }
    }
    void recycle(u64 id) {  // mod.rs:566-625
        u32 i = id_index(id);
        values[i] = T(0);
        memset(children[i], 0, 64);
        hashes[i] = 0;
        ref_counts[i] = 0;
        generations[i] += 1;
        if (generations[i] >= MAX_GENERATION) {
            generations[i] = 0;
            stats.generations_overflows += 1;
        }
        if (generations[i] > stats.max_generation) stats.max_generation = generations[i];
        free_indices.push_back(i);
        stats.alive_nodes -= 1;
        stats.total_deallocations += 1;
        stats.recycled_nodes += 1;
        if (id_is_leaf(id))
            stats.leaf_nodes -= 1;
        else
            stats.branch_nodes -= 1;
    }

    u64 get_or_create_leaf(T value) {  // mod.rs:627-710
        u64 hash = fx_leaf_hash(value);
        auto* s = patterns[1].find(hash, [&](u64 id) { return values[id_index(id)] == value; });
        if (s) {
            u64 id = s->id;
            if (!is_valid(id)) throw RefPanic("Expired node in patterns");  // :662-667
            inc_ref(id);
            stats.total_cache_hits += 1;
            stats.leaf_cache_hits += 1;
            return id;
        }
        u32 index = next_free_index();
        u64 id = id_leaf(index, generations[index]);
        patterns[1].insert(hash, id);
        values[index] = value;
        hashes[index] = hash;
        inc_ref(id);
        stats.leaf_nodes += 1;
        stats.patterns += 1;
        stats.total_cache_misses += 1;
        stats.leaf_cache_misses += 1;
        return id;
    }

    // get_or_create_branch — mod.rs:716-829.  Every non-empty child carries one bumped
    // reference: kept on a miss, released on a hit.
    u64 get_or_create_branch(const u64* c, u8 types, u8 mask) {
        u64 hash = fx_branch_hash(c, types, mask);
        auto* s = patterns[0].find(hash, [&](u64 id) {
            return id != ID_EMPTY && id_types(id) == types && id_mask(id) == mask &&
                   memcmp(children[id_index(id)], c, 64) == 0;
        });
        if (s) {
            u64 id = s->id;
            dec_child_refs(c);  // :767
            inc_ref(id);        // :772
            stats.total_cache_hits += 1;
            stats.branch_cache_hits += 1;
            return id;
        }
        u32 index = next_free_index();
        u64 id = id_branch(index, generations[index], types, mask);
        patterns[0].insert(hash, id);
        T cv[8];
        for (int i = 0; i < 8; ++i) cv[i] = values[id_index(c[i])];  // :794 (EMPTY -> slot 0 -> 0)
        memcpy(children[index], c, 64);
        values[index] = calc_average(cv);
        hashes[index] = hash;
        inc_ref(id);
        stats.branch_nodes += 1;
        stats.patterns += 1;
        stats.total_cache_misses += 1;
        stats.branch_cache_misses += 1;
        return id;
    }
};

// ---------------------------------------------------------------------------------
// Batch<T> — core/batch.rs:39-45 (struct), :63-81 (new), :145-175 (just_set),
// :178-184 (just_fill), :187-195 (just_clear).  The arrays are borrowed views so the
// oracle consumes byte-for-byte the same buffers as the CUDA path.
// ---------------------------------------------------------------------------------
inline size_t batch_blocks(int max_depth) {  // batch.rs:67-72
    int lower = max_depth > 0 ? max_depth - 1 : 0;
    return size_t(1) << (3 * lower);
}
template <class T>
struct BatchView {
    const u8* masks;   // [B][2] = (set_mask, clear_mask)
    const T* values;   // [B][8]
    bool has_fill;
    T fill;
    bool has_patches;
    int max_depth;
};
template <class T>
inline void batch_just_set(u8* masks, T* values, int x, int y, int z, T v) {  // batch.rs:145-175
    u32 full = encode_child_index_path(x, y, z);
    size_t pi = full >> 3;
    int idx = full & 7;
    u8 bit = u8(1u << idx);
    if (v != T(0)) {
        masks[2 * pi] |= bit;
        masks[2 * pi + 1] &= u8(~bit);
    } else {
        masks[2 * pi] &= u8(~bit);
        masks[2 * pi + 1] |= bit;
    }
    values[8 * pi + idx] = v;
}

// ---------------------------------------------------------------------------------
// set_batch_at_depth_iterative — spatial/voxtree.rs:724-1118 (called from
// set_batch_at_root :708-722 with initial depth 0).
// ---------------------------------------------------------------------------------
template <class T>
u64 set_batch_at_root(Interner<T>& in, u64 initial_node, int max_depth, const BatchView<T>& b) {
    if (initial_node == ID_INVALID) throw RefPanic("invalid root");  // :714
    // Phase 0 — fill (:742-754)
    if (b.has_fill) initial_node = in.get_or_create_leaf(b.fill);
    if (!b.has_patches) return initial_node;  // :756-758
    if (max_depth < 2) throw RefPanic("max_depth < 2 underflows target_depth - 1");  // :913,934

    const size_t data_len = batch_blocks(max_depth);
    std::vector<u64> cur(data_len, ID_INVALID), nxt(data_len, ID_INVALID);  // :772-773
    std::vector<size_t> paths, next_paths;
    paths.reserve(data_len);
    next_paths.reserve(data_len);

    // Phase 1 — blocks at depth max_depth-1 (:778-897)
    for (size_t pi = 0; pi < data_len; ++pi) {
        u8 set_mask = b.masks[2 * pi];
        if (set_mask == 0) continue;  // clear_mask is never read (:778-781)
        size_t path = pi << 3;
        u64 current = initial_node, leaf_node = ID_EMPTY;
        for (int cd = 0; cd < max_depth - 1; ++cd) {  // :792-822
            if (id_is_branch(current)) {
                int idx = int((path >> ((max_depth - cd - 1) * 3)) & 7);
                current = in.get_child_id(current, idx);
            } else {
                leaf_node = current;
                current = ID_EMPTY;
            }
            if (id_is_empty(current)) break;
        }
        const T* values = b.values + 8 * pi;
        bool all_same = set_mask == 0xFF;  // :826
        for (int i = 1; all_same && i < 8; ++i) all_same = values[i] == values[0];
        if (!all_same) {
            u64 children[8];
            u8 types, mask;
            if (!id_is_empty(current)) {  // :829-834 (panics if `current` is a leaf)
                memcpy(children, in.children[id_index(current)], 64);
                types = id_types(current);
                mask = id_mask(current);
            } else if (id_is_leaf(leaf_node)) {  // :835-839
                int n = __builtin_popcount(set_mask);
                in.inc_ref_by(leaf_node, u32(8 - n));
                for (int i = 0; i < 8; ++i) children[i] = leaf_node;
                types = mask = 0xFF;
            } else {  // :840-842
                memset(children, 0, 64);
                types = mask = 0;
            }
            u8 modified = 0;
            for (u8 bits = set_mask; bits;) {  // :846-863
                int idx = __builtin_ctz(bits);
                bits &= u8(~(1u << idx));
                T value = values[idx];
                if (!id_is_empty(children[idx]) && in.get_value(children[idx]) == value) continue;
                children[idx] = in.get_or_create_leaf(value);
                types |= u8(1u << idx);
                mask |= u8(1u << idx);
                modified |= u8(1u << idx);
            }
            if (modified == 0) continue;  // :865-868
            if (id_is_empty(leaf_node)) {  // :870-881
                for (u8 bits = u8(~modified); bits;) {
                    int idx = __builtin_ctz(bits);
                    bits &= u8(~(1u << idx));
                    if (!id_is_empty(children[idx])) in.inc_ref_by(children[idx], 1);
                }
            }
            cur[pi] = in.get_or_create_branch(children, types, mask);  // :883-886
            paths.push_back(path);
        } else {  // :887-896
            in.stats.collapsed_branches += 1;
            T first = values[__builtin_ctz(set_mask)];
            cur[pi] = in.get_or_create_leaf(first);
            paths.push_back(path);
        }
    }

    // Phase 2 — integrate upwards (:905-1106)
    if (paths.empty()) return ID_INVALID;  // :905-911
    int target_depth = max_depth - 1;
    while (!paths.empty()) {
        size_t path = paths.back();
        paths.pop_back();
        int pmd = target_depth > 1 ? target_depth - 2 : 0;  // :922-926
        size_t pmask = path_mask(max_depth, pmd);
        u64 equivalent = initial_node, leaf_id = ID_EMPTY;
        for (int cd = 0; cd < target_depth - 1; ++cd) {  // :934-946
            if (id_is_leaf(equivalent)) {
                leaf_id = equivalent;
                equivalent = ID_EMPTY;
                break;
            } else {
                int idx = int((path >> ((max_depth - cd - 1) * 3)) & 7);
                equivalent = in.get_child_id(equivalent, idx);
            }
            if (id_is_empty(equivalent)) break;
        }
        if (id_is_leaf(equivalent)) {  // :948-951
            leaf_id = equivalent;
            equivalent = ID_EMPTY;
        }
        u64 children[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        u8 types = 0, mask = 0;
        bool has_next = true;
        while (has_next) {  // :959-1000
            int tidx = int((path >> ((max_depth - target_depth) * 3)) & 7);
            bool have_next_path = !paths.empty();
            size_t next_path = have_next_path ? paths.back() : 0;
            if (target_depth == 1)
                has_next = have_next_path;
            else
                has_next = have_next_path && ((path & pmask) == (next_path & pmask));
            size_t cpi = path >> 3;
            u64 id = cur[cpi];
            children[tidx] = id;
            cur[cpi] = ID_INVALID;
            types |= u8(u8(id_is_leaf(id)) << tidx);
            mask |= u8(1u << tidx);
            if (has_next && !paths.empty()) {
                path = paths.back();
                paths.pop_back();
            }
        }
        u8 existing_mask = id_mask(equivalent);  // :1013 (equivalent is a branch or EMPTY here)
        u8 inv_mask = u8(~mask);
        u8 cloned = existing_mask & inv_mask;
        if (mask != 0xFF) {  // :1017-1048
            if (cloned != 0) {
                const u64* ex = in.children[id_index(equivalent)];
                for (u8 bits = cloned; bits;) {
                    int idx = __builtin_ctz(bits);
                    bits &= u8(~(1u << idx));
                    children[idx] = ex[idx];
                    types |= u8(u8(id_is_leaf(children[idx])) << idx);
                    mask |= u8(1u << idx);
                }
            } else if (!id_is_empty(leaf_id)) {
                u32 n = u32(__builtin_popcount(inv_mask));
                types |= inv_mask;
                mask |= inv_mask;
                for (u8 bits = inv_mask; bits;) {
                    int idx = __builtin_ctz(bits);
                    children[idx] = leaf_id;
                    bits &= u8(~(1u << idx));
                }
                in.inc_ref_by(leaf_id, n);
            }
        }
        bool all_same = types == 0xFF;  // :1050
        for (int i = 1; all_same && i < 8; ++i) all_same = children[i] == children[0];
        u64 new_id;
        if (!all_same) {  // :1052-1061
            for (u8 bits = cloned; bits;) {
                int idx = __builtin_ctz(bits);
                bits &= u8(~(1u << idx));
                in.inc_ref(children[idx]);
            }
            new_id = in.get_or_create_branch(children, types, mask);
        } else {  // :1062-1075
            in.stats.collapsed_branches += 1;
            u32 dec = cloned != 0 ? std::min<u32>(u32(__builtin_popcount(cloned)), 7u) : 7u;
            in.dec_ref_by(children[0], dec);
            new_id = children[0];
        }
        size_t np = path & pmask;  // :1083-1085
        next_paths.push_back(np);
        nxt[np >> 3] = new_id;
        if (paths.empty()) {  // :1092-1105
            cur.swap(nxt);
            paths.swap(next_paths);
            next_paths.clear();
            target_depth -= 1;
            if (target_depth == 0) break;
        }
    }
    return cur[paths[0] >> 3];  // :1108
}

// ---------------------------------------------------------------------------------
// VoxTree<T> — spatial/voxtree.rs:108-142 (struct/new), :264-292 (fill/clear),
// :303-328 (apply_batch), :144-160 (get).
// ---------------------------------------------------------------------------------
struct Tree {
    int max_depth;
    u64 root_id = ID_EMPTY;
    bool dirty = false;
    explicit Tree(int d) : max_depth(d) {}
};

template <class T>
bool tree_apply_batch(Interner<T>& in, Tree& t, const BatchView<T>& b) {  // :303-328
    u64 new_root = set_batch_at_root(in, t.root_id, t.max_depth, b);
    if (new_root != ID_INVALID) {
        if (!id_is_empty(t.root_id)) {
            if (new_root == t.root_id) throw RefPanic("assert_ne!(new_root_id, self.root_id)");  // :311
            in.dec_ref_recursive(t.root_id);
        }
        if (!in.is_valid(new_root)) throw RefPanic("Invalid new root id");  // :316-319
        t.root_id = new_root;
        t.dirty = true;
        return true;
    }
    return false;
}
template <class T>
void tree_clear(Interner<T>& in, Tree& t) {  // :283-292
    if (!id_is_empty(t.root_id)) {
        in.dec_ref_recursive(t.root_id);
        t.root_id = ID_EMPTY;
        t.dirty = true;
    }
}
template <class T>
void tree_fill(Interner<T>& in, Tree& t, T v) {  // :264-281
    if (v != T(0)) {
        if (!id_is_empty(t.root_id)) in.dec_ref_recursive(t.root_id);
        t.root_id = in.get_or_create_leaf(v);
        t.dirty = true;
    } else
        tree_clear(in, t);
}

// get_at_depth — utils/common.rs:122-156.  Returns false for None.
template <class T>
bool tree_get(const Interner<T>& in, const Tree& t, int x, int y, int z, T* out) {
    int md = t.max_depth;
    int n = 1 << md;
    if (x < 0 || x >= n || y < 0 || y >= n || z < 0 || z >= n)
        throw RefPanic("position out of bounds");  // voxtree.rs:146-148
    u64 node = t.root_id;
    int depth = 0;
    while (!id_is_empty(node)) {
        if (depth >= md) {
            T v = in.get_value(node);
            if (v != T(0)) {
                *out = v;
                return true;
            }
            return false;
        }
        if (id_is_branch(node)) {
            int shift = md - depth - 1;  // child_index_macro_2, common.rs:104-111
            int idx = ((x >> shift) & 1) | (((y >> shift) & 1) << 1) | (((z >> shift) & 1) << 2);
            node = in.get_child_id(node, idx);
            depth += 1;
        } else {
            *out = in.get_value(node);
            return true;
        }
    }
    return false;
}

// to_vec — utils/common.rs:158-246.  Dense layout index = y*N*N + z*N + x (:229-238).  `md` is the depth the
// caller asks for: MaxDepth::for_lod (core/max_depth.rs:137-140, used by world/voxchunk.rs:267) — branches
// standing at that depth contribute their LOD value.
template <class T>
void tree_to_vec(const Interner<T>& in, const Tree& t, T* data, int md = -1) {
    if (md < 0) md = t.max_depth;
    size_t n = size_t(1) << md;
    size_t size = n * n * n;
    if (!id_is_branch(t.root_id)) {  // :172-174
        T v = in.get_value(t.root_id);
        for (size_t i = 0; i < size; ++i) data[i] = v;
        return;
    }
    memset(data, 0, size * sizeof(T));
    if (id_is_empty(t.root_id)) return;
    struct Item {
        u64 id;
        int x, y, z, depth;
    };
    std::vector<Item> stack;
    stack.push_back({t.root_id, 0, 0, 0, 0});
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (id_is_branch(it.id) && it.depth < md) {
            int half = 1 << (md - it.depth - 1);
            const u64* ch = in.children[id_index(it.id)];
            for (int i = 7; i >= 0; --i)
                if (!id_is_empty(ch[i]))
                    stack.push_back({ch[i], it.x + (i & 1) * half, it.y + ((i >> 1) & 1) * half,
                                     it.z + ((i >> 2) & 1) * half, it.depth + 1});
        } else {
            T v = in.get_value(it.id);
            if (v != T(0)) {
                size_t side = size_t(1) << (md - it.depth);
                for (size_t y = it.y; y < it.y + side; ++y)
                    for (size_t z = it.z; z < it.z + side; ++z) {
                        T* row = data + y * n * n + z * n + it.x;
                        for (size_t k = 0; k < side; ++k) row[k] = v;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// VTM payload — VoxModel::serialize / deserialize (world/voxmodel.rs:177-294, :296-408),
// serialize_chunk / deserialize_chunk (world/voxchunk.rs:382-440), varints (io/varint.rs),
// interner side: deserialize_leaf / preallocate_branch_id / deserialize_branch
// (interner/mod.rs:930-1000), VoxTree::set_root_id (spatial/voxtree.rs:135-141).
// The reference walks `self.chunks` (a hash map, arbitrary order); here the chunks are
// written in the order given.
// ---------------------------------------------------------------------------------
inline void put_varint(std::vector<u8>& o, u64 v) {  // io/varint.rs:5-32 (u32 and usize forms agree)
    while (v >= 0x80) {
        o.push_back(u8((v & 0x7F) | 0x80));
        v >>= 7;
    }
    o.push_back(u8(v));
}
inline void put_be32(std::vector<u8>& o, u32 v) {
    for (int s = 24; s >= 0; s -= 8) o.push_back(u8(v >> s));
}
template <class T>
inline void put_value_be(std::vector<u8>& o, T v) {  // core/voxel.rs:26-28 (to_be_bytes)
    for (int s = int(sizeof(T)) * 8 - 8; s >= 0; s -= 8) o.push_back(u8(u64(typename std::make_unsigned<T>::type(v)) >> s));
}
inline const u8 VTC_MAGIC[12] = {'V', 'o', 'x', 'T', 'r', 'e', 'e', 'C', 'h', 'u', 'n', 'k'};  // io/consts.rs:3

template <class T>
void model_serialize(const Interner<T>& in, size_t n, const int32_t* pos, const u64* roots, std::vector<u8>& out) {
    std::vector<u64> leaves, branches;  // :189-195 every id the two pattern maps hold
    for (const auto& s : in.patterns[1].slots)
        if (s.id != ID_INVALID) leaves.push_back(s.id);
    for (const auto& s : in.patterns[0].slots)
        if (s.id != ID_INVALID) branches.push_back(s.id);
    auto by_index = [](u64 a, u64 b) { return id_index(a) < id_index(b); };
    std::sort(leaves.begin(), leaves.end(), by_index);      // :199-200
    std::sort(branches.begin(), branches.end(), by_index);
    std::vector<u32> id_map(in.next_index, 0);              // :186-187 id_map[0] = 0
    u32 next_id = 1;
    for (u64 id : leaves) id_map[id_index(id)] = next_id++;  // :202-205
    for (u64 id : branches)                                   // :207-215
        if (id_index(id) != 0) id_map[id_index(id)] = next_id++;
    put_be32(out, u32(leaves.size()));                        // :231
    for (u64 id : leaves) {                                   // :232-240
        put_varint(out, id_map[id_index(id)]);
        put_value_be<T>(out, in.values[id_index(id)]);
    }
    put_be32(out, u32(branches.size()) - 1);                  // :242 (slot 0, the empty branch, is not written)
    for (u64 id : branches) {                                 // :243-268
        if (id_index(id) == 0) continue;
        put_varint(out, id_map[id_index(id)]);
        out.push_back(id_mask(id));
        for (int k = 0; k < 8; ++k) {
            u64 ch = in.children[id_index(id)][k];
            if (id_is_empty(ch)) continue;
            put_varint(out, id_map[id_index(ch)]);
        }
        put_value_be<T>(out, in.values[id_index(id)]);
    }
    std::vector<u8> chunks;
    for (size_t c = 0; c < n; ++c) {                          // voxchunk.rs:382-405
        chunks.insert(chunks.end(), VTC_MAGIC, VTC_MAGIC + 12);
        for (int a = 0; a < 3; ++a) put_be32(chunks, u32(pos[3 * c + a]));
        put_varint(chunks, id_map[id_index(roots[c])]);
    }
    put_be32(out, u32(n));                                    // :281-284
    out.insert(out.end(), chunks.begin(), chunks.end());      // :286-288
}

struct VtmReader {
    const u8* p;
    const u8* end;
    u8 byte() {
        if (p >= end) throw RefPanic("unexpected end of VTM data");
        return *p++;
    }
    u32 be32() {
        u32 v = 0;
        for (int i = 0; i < 4; ++i) v = (v << 8) | byte();
        return v;
    }
    u32 varint() {  // io/varint.rs:62-96
        u32 r = 0;
        int shift = 0;
        for (;;) {
            u8 b = byte();
            r |= u32(b & 0x7F) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
            if (shift >= 32) throw RefPanic("varint too long");
        }
        return r;
    }
    template <class T>
    T value_be() {
        u64 v = 0;
        for (size_t i = 0; i < sizeof(T); ++i) v = (v << 8) | byte();
        return T(v);
    }
};

// Into a FRESH interner (the reference asserts that file id == next index, mod.rs:933,948).
// Returns the chunks' positions and root ids.
template <class T>
void model_deserialize(Interner<T>& in, const u8* data, size_t len, std::vector<int32_t>& pos, std::vector<u64>& roots) {
    VtmReader r{data, data + len};
    const u32 leaf_size = r.be32();  // :310
    std::vector<u64> by_file_id(1, ID_EMPTY);
    std::vector<u8> is_leaf(1, 0);
    auto place = [&](u32 id, u64 block, bool leaf) {
        if (id >= by_file_id.size()) {
            by_file_id.resize(id + 1, ID_INVALID);
            is_leaf.resize(id + 1, 0);
        }
        by_file_id[id] = block;
        is_leaf[id] = leaf;
    };
    for (u32 k = 0; k < leaf_size; ++k) {  // :317-326 + mod.rs:941-964
        u32 id = r.varint();
        T value = r.template value_be<T>();
        u32 index = in.next_free_index();
        if (index != id) throw RefPanic("Invalid block id");
        if (in.generations[index] != 0) throw RefPanic("Invalid generation");
        u64 block = id_leaf(index, 0);
        in.values[index] = value;
        in.hashes[index] = fx_leaf_hash(value);
        in.patterns[1].insert(in.hashes[index], block);  // (no stats besides the allocation's: mod.rs:941-964)
        place(id, block, true);
    }
    const u32 branch_size = r.be32();  // :328
    struct Rec {
        u32 id;
        u32 children[8];
        T lod;
    };
    std::vector<Rec> recs(branch_size);
    for (u32 k = 0; k < branch_size; ++k) {  // :335-364 + mod.rs:930-939
        Rec& rec = recs[k];
        rec.id = r.varint();
        if (rec.id == 0) throw RefPanic("branch id 0");
        u8 mask = r.byte(), types = 0;
        if (mask == 0) throw RefPanic("empty branch mask");
        for (int c = 0; c < 8; ++c) {
            rec.children[c] = 0;
            if (!(mask >> c & 1)) continue;
            rec.children[c] = r.varint();
            if (rec.children[c] < is_leaf.size() && is_leaf[rec.children[c]]) types |= u8(1u << c);
        }
        rec.lod = r.template value_be<T>();
        u32 index = in.next_free_index();
        if (index != rec.id) throw RefPanic("Invalid block id");
        if (in.generations[index] != 0) throw RefPanic("Invalid generation");
        place(rec.id, id_branch(index, 0, types, mask), false);
    }
    for (const Rec& rec : recs) {  // :366-394 + mod.rs:966-1000
        u64 block = by_file_id[rec.id];
        u64 ch[8];
        for (int c = 0; c < 8; ++c) ch[c] = (id_mask(block) >> c & 1) ? by_file_id[rec.children[c]] : ID_EMPTY;
        u32 index = id_index(block);
        memcpy(in.children[index], ch, 64);
        in.values[index] = rec.lod;
        in.hashes[index] = fx_branch_hash(ch, id_types(block), id_mask(block));
        in.patterns[0].insert(in.hashes[index], block);
        for (int c = 0; c < 8; ++c)
            if (!id_is_empty(ch[c])) in.inc_ref(ch[c]);  // inc_all_child_refs
    }
    const u32 n = r.be32();  // :410
    for (u32 c = 0; c < n; ++c) {  // voxchunk.rs:407-440
        for (int k = 0; k < 12; ++k)
            if (r.byte() != VTC_MAGIC[k]) throw RefPanic("bad chunk magic");
        for (int a = 0; a < 3; ++a) pos.push_back(int32_t(r.be32()));
        u32 root = r.varint();
        if (root >= by_file_id.size() || by_file_id[root] == ID_INVALID) throw RefPanic("unknown root id");
        u64 block = by_file_id[root];
        in.inc_ref(block);  // set_root_id, voxtree.rs:135-141 (an EMPTY root bumps slot 0, as in the reference)
        roots.push_back(block);
    }
}

// ---------------------------------------------------------------------------------
// Occupancy masks — the greedy mesher's input.  OccupancyDataBuilder (utils/mesh.rs:181-193,
// :247-285), fill_masks_for_region (:418-513), generate_occupancy_masks (:515-596).
// Planes (mesh.rs:50-67): MAX_VOXELS_PER_AXIS = 64, PLANE_SIZE = 64*64,
//   YZ at 0          word[y*64 + z], bit x
//   XZ at PLANE_SIZE word[z*64 + x], bit y
//   XY at 2*PLANE    word[y*64 + x], bit z
// `external` / `external_exists` belong to generate_external_occupancy_mask (:318-416), not restated.
// ---------------------------------------------------------------------------------
constexpr size_t OCC_AXIS = 64;                       // mesh.rs:50
constexpr size_t OCC_PLANE = OCC_AXIS * OCC_AXIS;     // :51
constexpr size_t OCC_ALL = OCC_PLANE * 3;             // :52
constexpr size_t OCC_YZ = 0, OCC_XZ = OCC_PLANE, OCC_XY = OCC_PLANE * 2;  // :65-67

struct OccupancyBuilder {  // mesh.rs:181-193, Default :247-261
    std::vector<u64> global = std::vector<u64>(OCC_ALL, 0);
    u64 global_active[6] = {0, 0, 0, 0, 0, 0};
    std::vector<std::pair<u64, std::vector<u64>>> per_material;  // HashMap<usize, Vec<u64>>
    std::vector<std::pair<u64, u64>> materials;                  // HashMap<usize, usize> (id -> voxel count)

    std::vector<u64>* find_masks(u64 id) {
        for (auto& e : per_material)
            if (e.first == id) return &e.second;
        return nullptr;
    }
    // build(): materials sorted by id, per_material in the same order — mesh.rs:263-285
    void sort_materials() {
        std::sort(materials.begin(), materials.end());
        std::sort(per_material.begin(), per_material.end(),
                  [](const auto& a, const auto& b) { return a.first < b.first; });
    }
};

// fill_masks_for_region — mesh.rs:418-513
inline void fill_masks_for_region(OccupancyBuilder& b, u32 ox, u32 oy, u32 oz, u32 side_u, u64 material_id) {
    const size_t side = side_u;
    const u64 volume = u64(side) * side * side;
    bool seen = false;  // :428-432
    for (auto& m : b.materials)
        if (m.first == material_id) {
            m.second += volume;
            seen = true;
        }
    if (!seen) b.materials.push_back({material_id, volume});
    if (side != OCC_AXIS) {  // :434
        std::vector<u64>* pm = b.find_masks(material_id);  // :435-438
        if (!pm) {
            b.per_material.push_back({material_id, std::vector<u64>(OCC_ALL, 0)});
            pm = &b.per_material.back().second;
        }
        if (ox + side > OCC_AXIS || oy + side > OCC_AXIS || oz + side > OCC_AXIS)
            throw RefPanic("region outside the 64^3 occupancy volume");  // Rust: shift overflow / index panic
        const u64 run = (u64(1) << side) - 1;  // :440
        const u64 x_mask = run << ox, y_mask = run << oy, z_mask = run << oz;
        b.global_active[0] |= y_mask;  // :451-461
        b.global_active[1] |= z_mask;
        b.global_active[2] |= z_mask;
        b.global_active[3] |= x_mask;
        b.global_active[4] |= y_mask;
        b.global_active[5] |= x_mask;
        for (size_t i = 0; i < side; ++i) {  // :463-486
            const size_t z = oz + i, y = oy + i;
            const size_t base_y = OCC_XZ + z * OCC_AXIS + ox;
            const size_t base_z = OCC_XY + y * OCC_AXIS + ox;
            const size_t base_x = OCC_YZ + y * OCC_AXIS + oz;
            for (size_t j = 0; j < side; ++j) {
                b.global[base_y + j] |= y_mask;
                (*pm)[base_y + j] |= y_mask;
                b.global[base_z + j] |= z_mask;
                (*pm)[base_z + j] |= z_mask;
                b.global[base_x + j] |= x_mask;
                (*pm)[base_x + j] |= x_mask;
            }
        }
    } else {  // :487-512 — one region covers the whole volume
        std::vector<u64>* pm = b.find_masks(material_id);
        if (pm)
            std::fill(pm->begin(), pm->end(), ~u64(0));
        else
            b.per_material.push_back({material_id, std::vector<u64>(OCC_ALL, ~u64(0))});
        std::fill(b.global.begin(), b.global.end(), ~u64(0));
        for (int k = 0; k < 6; ++k) b.global_active[k] = ~u64(0);
    }
}

// generate_occupancy_masks — mesh.rs:515-596.  `max_depth` is the depth the caller unfolds to
// (MaxDepth::for_lod at the call sites, world/voxchunk.rs); material_id = value as usize (core/voxel.rs:85-87).
template <class T>
void generate_occupancy_masks(const Interner<T>& in, OccupancyBuilder& b, u64 root_id, int max_depth, u32 offx,
                              u32 offy, u32 offz) {
    if (id_is_empty(root_id)) return;  // :530-537
    auto mat = [](T v) { return u64(int64_t(v)); };  // `*self as usize`: sign-extending for signed T
    if (!id_is_branch(root_id)) {  // :543-556
        T v = in.get_value(root_id);
        if (v != T(0)) fill_masks_for_region(b, offx, offy, offz, u32(1) << max_depth, mat(v));
        return;
    }
    struct Item {
        u64 id;
        u32 x, y, z, depth;
    };
    std::vector<Item> stack;  // :558-559
    stack.push_back({root_id, 0, 0, 0, 0});
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (id_is_branch(it.id) && it.depth < u32(max_depth)) {  // :562-579
            const u32 half = u32(1) << (max_depth - it.depth - 1);
            const u64* ch = in.children[id_index(it.id)];
            for (int i = 7; i >= 0; --i)
                if (!id_is_empty(ch[i]))
                    stack.push_back({ch[i], it.x + (u32(i) & 1) * half, it.y + ((u32(i) & 2) >> 1) * half,
                                     it.z + ((u32(i) & 4) >> 2) * half, it.depth + 1});
        } else {  // :580-589
            T v = in.get_value(it.id);
            if (v != T(0))
                fill_masks_for_region(b, offx + it.x, offy + it.y, offz + it.z, u32(1) << (max_depth - it.depth),
                                      mat(v));
        }
    }
}


// ---------------------------------------------------------------------------------
// Voxeliser — triangle -> voxel tests that fill a Batch (the step before the path).
//   voxelis-math/src/lib.rs:3-127     triangle_cube_intersection
//                          :129-151   point_in_or_on_cube
//                          :153-178   point_in_or_on_triangle
//                          :180-204   edge_quad_intersection      :206-214 point_in_quad
//   voxelis-voxelize/src/lib.rs:159-249   Voxelizer::voxelize_chunk
// All arithmetic is f64 in the reference's order of operations; the vector helpers restate glam 0.29.3
// (Cargo.lock; not vendored under /root/reference — "glam operation order: parity unpinned", pinned only through
// the known answers of the reference's own voxelis-math unit tests, tests/test_oracle_voxelize.py):
//   dot = x*x' + y*y' + z*z' (left to right), cross = (y*z' - y'*z, z*x' - z'*x, x*y' - x'*y),
//   length = sqrt(dot(v, v)), normalize = v * (1 / length).  Built with -ffp-contract=off (oracle/Makefile).
// ---------------------------------------------------------------------------------
struct DV3 {
    double x, y, z;
};
inline DV3 dv(double x, double y, double z) { return DV3{x, y, z}; }
inline DV3 operator+(DV3 a, DV3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline DV3 operator-(DV3 a, DV3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline DV3 operator*(DV3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline DV3 operator*(double s, DV3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline DV3 operator/(DV3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline DV3 vmin(DV3 a, DV3 b) { return {a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
inline DV3 vmax(DV3 a, DV3 b) { return {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }
inline double vdot(DV3 a, DV3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline DV3 vcross(DV3 a, DV3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline double vlength(DV3 a) { return std::sqrt(vdot(a, a)); }
inline DV3 vnormalize(DV3 a) { return a * (1.0 / vlength(a)); }
inline double f64_signum(double v) { return v != v ? v : (std::signbit(v) ? -1.0 : 1.0); }  // f64::signum

inline bool point_in_or_on_cube(DV3 p, DV3 cmin, DV3 cmax) {  // voxelis-math lib.rs:129-151
    const double cube_size = vlength(cmax - cmin);
    const double epsilon = cube_size * 1e-8;
    if (cube_size < 1e-8) return vlength(p - cmin) < epsilon;
    return p.x >= cmin.x - epsilon && p.x <= cmax.x + epsilon && p.y >= cmin.y - epsilon && p.y <= cmax.y + epsilon &&
           p.z >= cmin.z - epsilon && p.z <= cmax.z + epsilon;
}

inline bool point_in_or_on_triangle(DV3 p, DV3 a, DV3 b, DV3 c) {  // :153-178
    const DV3 v0 = b - a, v1 = c - a, v2 = p - a;
    const double dot00 = vdot(v0, v0), dot01 = vdot(v0, v1), dot02 = vdot(v0, v2), dot11 = vdot(v1, v1),
                 dot12 = vdot(v1, v2);
    const double denom = dot00 * dot11 - dot01 * dot01;
    if (std::fabs(denom) < 1e-8) return false;
    const double inv = 1.0 / denom;
    const double u = (dot11 * dot02 - dot01 * dot12) * inv;
    const double v = (dot00 * dot12 - dot01 * dot02) * inv;
    return u >= 0.0 && v >= 0.0 && (u + v) <= 1.0;
}

inline bool point_in_quad(DV3 p, const DV3* q) {  // :206-214
    return point_in_or_on_triangle(p, q[0], q[1], q[2]) || point_in_or_on_triangle(p, q[0], q[2], q[3]);
}

inline bool edge_quad_intersection(DV3 e1, DV3 e2, const DV3* q) {  // :180-204
    const DV3 normal = vnormalize(vcross(q[1] - q[0], q[2] - q[0]));
    const double denom = vdot(normal, e2 - e1);
    if (std::fabs(denom) < 1e-8) return false;
    const double t = vdot(normal, q[0] - e1) / denom;
    if (!(t >= 0.0 && t <= 1.0)) return false;
    return point_in_quad(e1 + t * (e2 - e1), q);
}

inline bool triangle_cube_intersection(DV3 tv0, DV3 tv1, DV3 tv2, DV3 cmin, DV3 cmax) {  // :3-127
    const DV3 tri_min = vmin(vmin(tv0, tv1), tv2), tri_max = vmax(vmax(tv0, tv1), tv2);
    const double epsilon = 1e-5;
    if (tri_max.x < cmin.x - epsilon || tri_min.x > cmax.x + epsilon || tri_max.y < cmin.y - epsilon ||
        tri_min.y > cmax.y + epsilon || tri_max.z < cmin.z - epsilon || tri_min.z > cmax.z + epsilon)
        return false;
    const DV3 normal = vcross(tv1 - tv0, tv2 - tv0);
    const double d = -vdot(normal, tv0);
    const DV3 cp[8] = {dv(cmin.x, cmin.y, cmin.z), dv(cmax.x, cmin.y, cmin.z), dv(cmax.x, cmax.y, cmin.z),
                       dv(cmin.x, cmax.y, cmin.z), dv(cmin.x, cmin.y, cmax.z), dv(cmax.x, cmin.y, cmax.z),
                       dv(cmax.x, cmax.y, cmax.z), dv(cmin.x, cmax.y, cmax.z)};
    const double sign = f64_signum(vdot(normal, cp[0]) + d);
    for (int i = 1; i < 8; ++i) {  // :41-50
        const double s = vdot(normal, cp[i]) + d;
        const double new_sign = f64_signum(s);
        if (std::fabs(s) < epsilon) continue;
        if (new_sign != sign) return true;
    }
    if (point_in_or_on_cube(tv0, cmin, cmax) || point_in_or_on_cube(tv1, cmin, cmax) ||
        point_in_or_on_cube(tv2, cmin, cmax))
        return true;  // :53-58
    for (int i = 0; i < 8; ++i)
        if (point_in_or_on_triangle(cp[i], tv0, tv1, tv2)) return true;  // :60-64
    const DV3 edges[3][2] = {{tv0, tv1}, {tv1, tv2}, {tv2, tv0}};  // :67
    const int fq[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {0, 3, 7, 4}, {1, 2, 6, 5}};  // :68-111
    for (int e = 0; e < 3; ++e)
        for (int f = 0; f < 6; ++f) {
            const DV3 q[4] = {cp[fq[f][0]], cp[fq[f][1]], cp[fq[f][2]], cp[fq[f][3]]};
            if (edge_quad_intersection(edges[e][0], edges[e][1], q)) return true;
        }
    return false;
}

inline int f64_as_i32(double v) {  // Rust `as i32`: saturating, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return int(-2147483647 - 1);
    return int(v);
}

// Voxelizer::voxelize_chunk — voxelis-voxelize/src/lib.rs:159-249.  faces[n][3] are 1-based vertex indices.
// Returns batch.has_patches(); masks / values are the Batch<T> arrays (zeroed by the caller = Batch::new).
template <class T>
bool voxelize_chunk(const int* chunk_position, int depth, double chunk_world_size, const double* mesh_min_,
                    size_t nfaces, const int32_t* faces, const double* vertices, u8* masks, T* values) {
    const int vpa = 1 << depth;
    const double voxel_size = chunk_world_size / double(vpa);  // voxelize_mesh :262
    const double epsilon = voxel_size * 1e-7;
    const DV3 splat = dv(epsilon, epsilon, epsilon);
    const DV3 mesh_min = dv(mesh_min_[0], mesh_min_[1], mesh_min_[2]);
    const DV3 cw_min = dv(double(chunk_position[0]), double(chunk_position[1]), double(chunk_position[2])) * chunk_world_size;
    const DV3 cw_max = cw_min + dv(chunk_world_size, chunk_world_size, chunk_world_size);
    bool has_patches = false;
    auto vert = [&](int32_t i) { return dv(vertices[3 * size_t(i - 1)], vertices[3 * size_t(i - 1) + 1], vertices[3 * size_t(i - 1) + 2]); };
    auto clampi = [&](int v) { return v < 0 ? 0 : v > vpa - 1 ? vpa - 1 : v; };
    for (size_t f = 0; f < nfaces; ++f) {
        const DV3 v1 = vert(faces[3 * f]) - mesh_min, v2 = vert(faces[3 * f + 1]) - mesh_min, v3 = vert(faces[3 * f + 2]) - mesh_min;
        const DV3 face_min = vmin(vmin(v1, v2), v3), face_max = vmax(vmax(v1, v2), v3);
        const DV3 omin = vmax(face_min, cw_min) - splat, omax = vmin(face_max, cw_max) + splat;
        if (omin.x >= omax.x || omin.y >= omax.y || omin.z >= omax.z) continue;  // :200-206
        const DV3 lo = (omin - cw_min) / voxel_size, hi = (omax - cw_min) / voxel_size;
        const int x0 = clampi(f64_as_i32(std::floor(lo.x))), y0 = clampi(f64_as_i32(std::floor(lo.y))), z0 = clampi(f64_as_i32(std::floor(lo.z)));
        const int x1 = clampi(f64_as_i32(std::ceil(hi.x))), y1 = clampi(f64_as_i32(std::ceil(hi.y))), z1 = clampi(f64_as_i32(std::ceil(hi.z)));
        for (int y = y0; y <= y1; ++y)
            for (int z = z0; z <= z1; ++z)
                for (int x = x0; x <= x1; ++x) {
                    const DV3 wp = cw_min + dv(double(x), double(y), double(z)) * voxel_size;  // :227-228
                    const DV3 wmin = wp - splat;
                    const DV3 wmax = wp + dv(voxel_size, voxel_size, voxel_size) + splat;
                    if (triangle_cube_intersection(v1, v2, v3, wmin, wmax)) {
                        batch_just_set<T>(masks, values, x, y, z, T(1));  // :239
                        has_patches = true;
                    }
                }
    }
    return has_patches;
}

}  // namespace vxo

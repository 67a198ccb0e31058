// oracle_capi.cpp — C entry points over voxelis_oracle.hpp for ctypes (tests/, bench.py's
// cpu_baseline leg, __graft_entry__.smoke()).  TEST INFRASTRUCTURE — never linked into or
// loaded by the product library (voxelis_b200/libvoxelis_b200.so).
//
// dtype codes match include/voxelis_b200.h: 0 = u8, 1 = i32.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>
#include <unordered_map>

#include "voxelis_oracle.hpp"

using namespace vxo;

namespace {
thread_local std::string g_err;

struct AnyInterner {
    int dtype;
    Interner<u8>* i8 = nullptr;
    Interner<int32_t>* i32 = nullptr;
    ~AnyInterner() {
        delete i8;
        delete i32;
    }
};

template <class F>
int guarded(F f) {
    try {
        return f();
    } catch (const RefPanic& e) {
        g_err = std::string("reference panic: ") + e.what();
        return -2;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

template <class T>
BatchView<T> view(const u8* masks, const void* values, int has_fill, int64_t fill, int has_patches,
                  int depth) {
    return BatchView<T>{masks, (const T*)values, has_fill != 0, T(fill), has_patches != 0, depth};
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_interner_create(size_t budget, int dtype) {
    AnyInterner* a = new AnyInterner{dtype};
    int rc = guarded([&] {
        if (dtype == 0)
            a->i8 = new Interner<u8>(budget);
        else if (dtype == 1)
            a->i32 = new Interner<int32_t>(budget);
        else
            throw std::runtime_error("bad dtype");
        return 0;
    });
    if (rc != 0) {
        delete a;
        return nullptr;
    }
    return a;
}
void orc_interner_destroy(void* h) { delete (AnyInterner*)h; }

void* orc_tree_create(int depth) {
    // MaxDepth::new asserts max < 7 (core/max_depth.rs:77-83); depth 7 is this build's
    // extension and is accepted here with the generalised PATH_MASKS.
    if (depth < 0 || depth > 7) {
        g_err = "Max depth exceeds allowed limit";
        return nullptr;
    }
    return new Tree(depth);
}
void orc_tree_destroy(void* t) { delete (Tree*)t; }
uint64_t orc_tree_root(void* t) { return ((Tree*)t)->root_id; }
// the tree takes over a root whose reference already exists (roots returned by orc_model_deserialize)
void orc_tree_adopt_root(void* t, uint64_t root) {
    ((Tree*)t)->root_id = root;
    ((Tree*)t)->dirty = true;
}
int orc_tree_dirty(void* t) { return ((Tree*)t)->dirty; }

size_t orc_batch_blocks(int depth) { return batch_blocks(depth); }

// Batch::just_set on caller-owned arrays (core/batch.rs:145-175).
void orc_batch_set(uint8_t* masks, void* values, int dtype, int x, int y, int z, int64_t v) {
    if (dtype == 0)
        batch_just_set<u8>(masks, (u8*)values, x, y, z, u8(v));
    else
        batch_just_set<int32_t>(masks, (int32_t*)values, x, y, z, int32_t(v));
}
uint32_t orc_encode_child_index_path(int x, int y, int z) { return encode_child_index_path(x, y, z); }
uint32_t orc_path_mask(int max_depth, int level) { return path_mask(max_depth, level); }
uint64_t orc_id_pack(uint32_t index, uint16_t gen, uint8_t types, uint8_t mask, int leaf) {
    uint64_t r = 0;
    guarded([&] {
        r = id_pack(index, gen, types, mask, leaf != 0);
        return 0;
    });
    return r;
}

// -> 1 changed, 0 unchanged, -2 the reference would panic, -1 other error
int orc_tree_apply_batch(void* ih, void* th, const uint8_t* masks, const void* values, int has_fill,
                         int64_t fill, int has_patches) {
    AnyInterner* a = (AnyInterner*)ih;
    Tree* t = (Tree*)th;
    return guarded([&] {
        if (a->dtype == 0)
            return int(tree_apply_batch(*a->i8, *t,
                                        view<u8>(masks, values, has_fill, fill, has_patches, t->max_depth)));
        return int(tree_apply_batch(*a->i32, *t,
                                    view<int32_t>(masks, values, has_fill, fill, has_patches, t->max_depth)));
    });
}

// Serial multi-chunk driver: the loop of voxelis-voxelize/src/lib.rs:357-361.  Batches are
// one contiguous slab [n][B] of masks and values; fills[n] (<0 = none... use has_fill[n]).
// roots_out[n], changed_out[n].  Trees are fresh (root EMPTY).
int orc_apply_batches_fresh(void* ih, int depth, size_t n, const uint8_t* masks, const void* values,
                            const uint8_t* has_fill, const int64_t* fills, const uint8_t* has_patches,
                            uint64_t* roots_out, uint8_t* changed_out) {
    AnyInterner* a = (AnyInterner*)ih;
    size_t B = batch_blocks(depth);
    return guarded([&] {
        for (size_t c = 0; c < n; ++c) {
            Tree t(depth);
            bool ch;
            int hf = has_fill ? has_fill[c] : 0;
            int64_t fv = fills ? fills[c] : 0;
            int hp = has_patches ? has_patches[c] : 1;
            if (a->dtype == 0)
                ch = tree_apply_batch(*a->i8, t,
                                      view<u8>(masks + c * B * 2, (const u8*)values + c * B * 8, hf, fv, hp, depth));
            else
                ch = tree_apply_batch(
                    *a->i32, t,
                    view<int32_t>(masks + c * B * 2, (const int32_t*)values + c * B * 8, hf, fv, hp, depth));
            roots_out[c] = t.root_id;
            if (changed_out) changed_out[c] = ch;
        }
        return 0;
    });
}

int orc_tree_fill(void* ih, void* th, int64_t v) {
    AnyInterner* a = (AnyInterner*)ih;
    return guarded([&] {
        if (a->dtype == 0)
            tree_fill(*a->i8, *(Tree*)th, u8(v));
        else
            tree_fill(*a->i32, *(Tree*)th, int32_t(v));
        return 0;
    });
}
int orc_tree_clear(void* ih, void* th) {
    AnyInterner* a = (AnyInterner*)ih;
    return guarded([&] {
        if (a->dtype == 0)
            tree_clear(*a->i8, *(Tree*)th);
        else
            tree_clear(*a->i32, *(Tree*)th);
        return 0;
    });
}
// -> 1 Some(*out), 0 None, <0 error
int orc_tree_get(void* ih, void* th, int x, int y, int z, int64_t* out) {
    AnyInterner* a = (AnyInterner*)ih;
    return guarded([&] {
        if (a->dtype == 0) {
            u8 v = 0;
            bool s = tree_get(*a->i8, *(Tree*)th, x, y, z, &v);
            *out = v;
            return int(s);
        }
        int32_t v = 0;
        bool s = tree_get(*a->i32, *(Tree*)th, x, y, z, &v);
        *out = v;
        return int(s);
    });
}
int orc_tree_to_vec(void* ih, void* th, void* dense) {
    AnyInterner* a = (AnyInterner*)ih;
    return guarded([&] {
        if (a->dtype == 0)
            tree_to_vec(*a->i8, *(Tree*)th, (u8*)dense);
        else
            tree_to_vec(*a->i32, *(Tree*)th, (int32_t*)dense);
        return 0;
    });
}
// to_vec from a bare root id (for roots returned by orc_apply_batches_fresh)
int orc_root_to_vec(void* ih, uint64_t root, int depth, void* dense) {
    Tree t(depth);
    t.root_id = root;
    return orc_tree_to_vec(ih, &t, dense);
}

// to_vec at a level of detail: depth - lod levels are unfolded (saturating, core/max_depth.rs:137-140)
int orc_root_to_vec_lod(void* ih, uint64_t root, int depth, int lod, void* dense) {
    AnyInterner* a = (AnyInterner*)ih;
    Tree t(depth);
    t.root_id = root;
    const int md = depth > lod ? depth - lod : 0;
    return guarded([&] {
        if (a->dtype == 0)
            tree_to_vec(*a->i8, t, (u8*)dense, md);
        else
            tree_to_vec(*a->i32, t, (int32_t*)dense, md);
        return 0;
    });
}

// generate_occupancy_masks (utils/mesh.rs:515-596) of n roots into ONE OccupancyDataBuilder, then build()'s
// ordering (:263-285).  `depth` = the depth unfolded (MaxDepth::for_lod); offsets[n][3] in voxels of that depth.
// Returns the number of materials; ids / counts / per_material (max_mat x 12288 words) are filled in id order.
long long orc_occupancy_masks(void* ih, size_t n, const uint64_t* roots, int depth, const uint32_t* offsets,
                              uint64_t* global, uint64_t* active, size_t max_mat, uint64_t* mat_ids,
                              uint64_t* mat_counts, uint64_t* per_material) {
    AnyInterner* a = (AnyInterner*)ih;
    long long out = 0;
    int rc = guarded([&] {
        OccupancyBuilder b;
        for (size_t i = 0; i < n; ++i) {
            const uint32_t* o = offsets + 3 * i;
            if (a->dtype == 0)
                generate_occupancy_masks(*a->i8, b, roots[i], depth, o[0], o[1], o[2]);
            else
                generate_occupancy_masks(*a->i32, b, roots[i], depth, o[0], o[1], o[2]);
        }
        b.sort_materials();
        if (b.materials.size() > max_mat) throw std::runtime_error("more materials than max_mat");
        memcpy(global, b.global.data(), OCC_ALL * 8);
        memcpy(active, b.global_active, 48);
        for (size_t m = 0; m < b.materials.size(); ++m) {
            mat_ids[m] = b.materials[m].first;
            mat_counts[m] = b.materials[m].second;
            memcpy(per_material + m * OCC_ALL, b.per_material[m].second.data(), OCC_ALL * 8);
        }
        out = (long long)b.materials.size();
        return 0;
    });
    return rc < 0 ? rc : out;
}

// voxelis-math known-answer entry points (tests/test_oracle_voxelize.py ports the crate's unit tests)
int orc_point_in_or_on_cube(const double* p, const double* cube) {
    return point_in_or_on_cube(dv(p[0], p[1], p[2]), dv(cube[0], cube[1], cube[2]), dv(cube[3], cube[4], cube[5]));
}
int orc_point_in_or_on_triangle(const double* p, const double* t) {
    return point_in_or_on_triangle(dv(p[0], p[1], p[2]), dv(t[0], t[1], t[2]), dv(t[3], t[4], t[5]), dv(t[6], t[7], t[8]));
}
int orc_edge_quad_intersection(const double* e, const double* q) {
    const DV3 quad[4] = {dv(q[0], q[1], q[2]), dv(q[3], q[4], q[5]), dv(q[6], q[7], q[8]), dv(q[9], q[10], q[11])};
    return edge_quad_intersection(dv(e[0], e[1], e[2]), dv(e[3], e[4], e[5]), quad);
}
int orc_point_in_quad(const double* p, const double* q) {
    const DV3 quad[4] = {dv(q[0], q[1], q[2]), dv(q[3], q[4], q[5]), dv(q[6], q[7], q[8]), dv(q[9], q[10], q[11])};
    return point_in_quad(dv(p[0], p[1], p[2]), quad);
}
int orc_triangle_cube_intersection(const double* t, const double* cube) {
    return triangle_cube_intersection(dv(t[0], t[1], t[2]), dv(t[3], t[4], t[5]), dv(t[6], t[7], t[8]),
                                      dv(cube[0], cube[1], cube[2]), dv(cube[3], cube[4], cube[5]));
}
// Voxelizer::voxelize_chunk (voxelis-voxelize/src/lib.rs:159-249) into zeroed Batch arrays; returns has_patches
int orc_voxelize_chunk(int dtype, const int* chunk_position, int depth, double chunk_world_size,
                       const double* mesh_min, size_t nfaces, const int32_t* faces, const double* vertices,
                       uint8_t* masks, void* values) {
    return guarded([&] {
        return int(dtype == 0 ? voxelize_chunk<u8>(chunk_position, depth, chunk_world_size, mesh_min, nfaces, faces,
                                                    vertices, masks, (u8*)values)
                              : voxelize_chunk<int32_t>(chunk_position, depth, chunk_world_size, mesh_min, nfaces, faces,
                                                         vertices, masks, (int32_t*)values));
    });
}

uint32_t orc_interner_ref(void* ih, uint64_t id) {
    AnyInterner* a = (AnyInterner*)ih;
    return a->dtype == 0 ? a->i8->get_ref(id) : a->i32->get_ref(id);
}
uint32_t orc_interner_next_index(void* ih) {
    AnyInterner* a = (AnyInterner*)ih;
    return a->dtype == 0 ? a->i8->next_index : a->i32->next_index;
}
size_t orc_interner_free_count(void* ih) {
    AnyInterner* a = (AnyInterner*)ih;
    return a->dtype == 0 ? a->i8->free_indices.size() : a->i32->free_indices.size();
}
size_t orc_interner_capacity(void* ih) {
    AnyInterner* a = (AnyInterner*)ih;
    return a->dtype == 0 ? a->i8->capacity : a->i32->capacity;
}
// 25 u64 counters in interner/stats.rs field order.
void orc_interner_stats(void* ih, uint64_t* out) {
    AnyInterner* a = (AnyInterner*)ih;
    const Stats& s = a->dtype == 0 ? a->i8->stats : a->i32->stats;
    memcpy(out, &s, sizeof(Stats));
}
// Copies the first next_index entries of each pool.  values_out is int64 per node.
void orc_interner_download(void* ih, uint64_t* children, int64_t* values, uint32_t* refs, uint16_t* gens) {
    AnyInterner* a = (AnyInterner*)ih;
    auto dl = [&](auto& in) {
        size_t n = in.next_index;
        if (children) memcpy(children, in.children, n * 64);
        if (values)
            for (size_t i = 0; i < n; ++i) values[i] = int64_t(in.values[i]);
        if (refs) memcpy(refs, in.ref_counts, n * 4);
        if (gens) memcpy(gens, in.generations, n * 2);
    };
    if (a->dtype == 0)
        dl(*a->i8);
    else
        dl(*a->i32);
}

// ---------------------------------------------------------------------------------
// DAG checker (works on downloaded pools, so the SAME code judges the oracle and the GPU
// interner).  Walks the DAG from roots[0..m) depth-first, children in index order, and
// numbers every distinct node at first completion (post-order).  The record stream
//     leaf   : [1, value, 0,0,0,0,0,0,0]
//     branch : [0, n(c0) .. n(c7)]      n(EMPTY) = 0, numbers start at 1
// followed by the root numbers is identical for two interners iff the rooted ordered DAGs
// are isomorphic (node ids permuted).  Outputs:
//   sig[2]            128-bit FNV-style digest of the stream
//   per_depth[2*(depth+1)]  (#distinct branch ids, #distinct leaf ids) reachable at each depth
//   totals[2]         distinct (branches, leaves) reachable
//   stream/stream_cap optional copy of the stream (u64 words); returns words needed
//   indeg             optional [n_nodes] in-degree from distinct reachable branches + roots
//   numbers           optional [n_nodes] canonical number of each node (0 = unreachable)
// Returns number of stream words, or -1 on a malformed pool (index out of range / cycle).
// ---------------------------------------------------------------------------------
long long orc_dag_signature(const uint64_t* children, const int64_t* values, size_t n_nodes,
                            const uint64_t* roots, size_t m, int depth, uint64_t* sig,
                            uint64_t* per_depth, uint64_t* totals, uint64_t* stream, size_t stream_cap,
                            uint32_t* indeg, uint32_t* numbers) {
    std::vector<uint32_t> number(n_nodes, 0);
    std::vector<uint8_t> state(n_nodes, 0);  // 0 new, 1 open, 2 done
    uint64_t h0 = 0xcbf29ce484222325ull, h1 = 0x84222325cbf29ce4ull;
    size_t words = 0;
    auto emit = [&](uint64_t w) {
        h0 = (h0 ^ w) * 0x100000001b3ull;
        h0 ^= h0 >> 29;
        h1 = (h1 + w) * 0x9E3779B97F4A7C15ull;
        h1 ^= h1 >> 32;
        if (stream && words < stream_cap) stream[words] = w;
        ++words;
    };
    uint32_t next_no = 1;
    uint64_t nb = 0, nl = 0;
    if (indeg) std::fill(indeg, indeg + n_nodes, 0u);
    struct Frame {
        uint64_t id;
        int k;
    };
    std::vector<Frame> st;
    for (size_t r = 0; r < m; ++r) {
        uint64_t root = roots[r];
        if (indeg && root != 0 && id_index(root) < n_nodes) indeg[id_index(root)] += 1;
        if (root == 0) continue;
        if (id_index(root) >= n_nodes) return -1;
        if (state[id_index(root)] == 2) continue;
        st.push_back({root, 0});
        while (!st.empty()) {
            Frame& f = st.back();
            uint32_t idx = id_index(f.id);
            if (id_is_leaf(f.id)) {
                if (state[idx] != 2) {
                    state[idx] = 2;
                    number[idx] = next_no++;
                    ++nl;
                    emit(1);
                    emit(uint64_t(values[idx]));
                    for (int i = 0; i < 7; ++i) emit(0);
                }
                st.pop_back();
                continue;
            }
            if (f.k == 0) {
                if (state[idx] == 2) {
                    st.pop_back();
                    continue;
                }
                state[idx] = 1;
            }
            bool pushed = false;
            while (f.k < 8) {
                uint64_t ch = children[size_t(idx) * 8 + f.k];
                f.k++;
                if (ch == 0) continue;
                uint32_t ci = id_index(ch);
                if (ci >= n_nodes) return -1;
                if (state[ci] == 1) return -1;  // cycle
                if (state[ci] == 2) continue;
                st.push_back({ch, 0});
                pushed = true;
                break;
            }
            if (pushed) continue;
            // all children done
            state[idx] = 2;
            number[idx] = next_no++;
            ++nb;
            emit(0);
            for (int i = 0; i < 8; ++i) {
                uint64_t ch = children[size_t(idx) * 8 + i];
                emit(ch == 0 ? 0 : number[id_index(ch)]);
                if (indeg && ch != 0) indeg[id_index(ch)] += 1;
            }
            st.pop_back();
        }
    }
    for (size_t r = 0; r < m; ++r) emit(roots[r] == 0 ? 0 : number[id_index(roots[r])]);
    if (numbers) memcpy(numbers, number.data(), n_nodes * sizeof(uint32_t));
    if (sig) {
        sig[0] = h0;
        sig[1] = h1;
    }
    if (totals) {
        totals[0] = nb;
        totals[1] = nl;
    }
    if (per_depth) {
        std::vector<uint64_t> level;
        for (size_t r = 0; r < m; ++r)
            if (roots[r] != 0) level.push_back(roots[r]);
        for (int d = 0; d <= depth; ++d) {
            std::sort(level.begin(), level.end());
            level.erase(std::unique(level.begin(), level.end()), level.end());
            uint64_t b = 0, l = 0;
            std::vector<uint64_t> next;
            for (uint64_t id : level) {
                if (id_is_leaf(id))
                    ++l;
                else {
                    ++b;
                    for (int i = 0; i < 8; ++i) {
                        uint64_t ch = children[size_t(id_index(id)) * 8 + i];
                        if (ch != 0) next.push_back(ch);
                    }
                }
            }
            per_depth[2 * d] = b;
            per_depth[2 * d + 1] = l;
            level.swap(next);
        }
    }
    return (long long)words;
}

// ---------------------------------------------------------------------------------
// CPU baseline timing helpers (bench.py cpu_baseline / --impl reference).  Fresh-tree
// apply over a slab of pre-built batches; one private interner per thread (the CPU analogue
// of per-GPU interners).  Returns elapsed seconds for the apply loop only.
// ---------------------------------------------------------------------------------
double orc_time_apply_fresh(int dtype, int depth, size_t budget, size_t n, const uint8_t* masks,
                            const void* values, int threads, uint64_t* roots_out) {
    size_t B = batch_blocks(depth);
    if (threads < 1) threads = 1;
    std::vector<void*> interners(threads, nullptr);
    for (int t = 0; t < threads; ++t) {
        interners[t] = orc_interner_create(budget, dtype);
        if (!interners[t]) return -1.0;
    }
    std::vector<uint64_t> roots(n);
    std::vector<int> rcs(threads, 0);
    size_t vsz = dtype == 0 ? 1 : 4;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        pool.emplace_back([&, t, lo, hi] {
            rcs[t] = orc_apply_batches_fresh(interners[t], depth, hi - lo, masks + lo * B * 2,
                                             (const uint8_t*)values + lo * B * 8 * vsz, nullptr, nullptr,
                                             nullptr, roots.data() + lo, nullptr);
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t) orc_interner_destroy(interners[t]);
    for (int t = 0; t < threads; ++t)
        if (rcs[t] != 0) return -1.0;
    if (roots_out) memcpy(roots_out, roots.data(), n * 8);
    return std::chrono::duration<double>(t1 - t0).count();
}

// VoxModel::serialize (world/voxmodel.rs:177-294): the VTM payload of n chunks (positions[n][3], roots[n]).
// Returns the payload size; writes it when it fits in `cap`.
int64_t orc_model_serialize(void* ih, size_t n, const int32_t* positions, const uint64_t* roots, uint8_t* out, size_t cap) {
    AnyInterner* a = (AnyInterner*)ih;
    std::vector<u8> buf;
    int rc = guarded([&] {
        if (a->dtype == 0)
            model_serialize<u8>(*a->i8, n, positions, roots, buf);
        else
            model_serialize<int32_t>(*a->i32, n, positions, roots, buf);
        return 0;
    });
    if (rc != 0) return rc;
    if (out && buf.size() <= cap) memcpy(out, buf.data(), buf.size());
    return int64_t(buf.size());
}
// VoxModel::deserialize (world/voxmodel.rs:296-408) into a FRESH interner.  Returns the number of chunks;
// positions_out[cap][3] / roots_out[cap] receive them.
int64_t orc_model_deserialize(void* ih, const uint8_t* data, size_t len, int32_t* positions_out, uint64_t* roots_out,
                              size_t cap) {
    AnyInterner* a = (AnyInterner*)ih;
    std::vector<int32_t> pos;
    std::vector<u64> roots;
    int rc = guarded([&] {
        if (a->dtype == 0)
            model_deserialize<u8>(*a->i8, data, len, pos, roots);
        else
            model_deserialize<int32_t>(*a->i32, data, len, pos, roots);
        return 0;
    });
    if (rc != 0) return rc;
    if (roots.size() <= cap) {
        if (positions_out) memcpy(positions_out, pos.data(), pos.size() * 4);
        if (roots_out) memcpy(roots_out, roots.data(), roots.size() * 8);
    }
    return int64_t(roots.size());
}

}  // extern "C"

/* voxelis_b200.h — C ABI of the B200-native batched SVO-DAG build path.
 *
 * Drop-in boundary for ONE path of WildPixelGames/voxelis v25.4.0:
 *     Batch  ->  VoxTree::apply_batch  ->  VoxInterner   (+ a new multi-chunk apply_batches)
 * The reference is a Rust library with no FFI surface of its own, so each entry point below is
 * shaped 1:1 on the Rust method it replaces (cited as file:line under
 * /root/reference/voxelis/src).  INTEGRATION.md shows the `extern "C"` block a Rust maintainer
 * would add on top of this header.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA or torch types in any signature (streams are passed
 *     as `void*`, i.e. a cudaStream_t, and may be NULL for the interner's own stream);
 *   - every function that can fail returns a negative vx_status; vx_last_error() gives the
 *     message for the calling thread.  The reference panics at these sites; this library never
 *     aborts the process;
 *   - threading follows the reference (`&mut VoxInterner`, world/voxmodel.rs:32): one mutator per
 *     interner at a time; reads may run concurrently with each other but not with an apply;
 *   - there is NO CPU fallback: with no usable CUDA device every compute entry point fails with
 *     VX_E_CUDA.
 */
#ifndef VOXELIS_B200_H
#define VOXELIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_ABI_VERSION 1

/* VoxelTrait numeric impls in scope (core/voxel.rs:77,93): u8 and i32. */
typedef enum vx_dtype { VX_U8 = 0, VX_I32 = 1 } vx_dtype;

typedef enum vx_status {
    VX_OK = 0,
    VX_E_INVALID = -1,      /* bad argument / handle                                            */
    VX_E_OOM = -2,          /* "Out of memory" — interner/macros.rs:38 (node capacity exhausted) */
    VX_E_CUDA = -3,         /* CUDA runtime error or no device                                   */
    VX_E_UNSUPPORTED = -4,  /* valid in the reference, not implemented on this path yet          */
    VX_E_BOUNDS = -5,       /* position out of range — spatial/voxtree.rs:146-148                */
    VX_E_BUDGET = -6,       /* "Requested budget is too small/large" — interner/mod.rs:63-68     */
    VX_E_POISONED = -7      /* a previous apply overflowed the interner; reset or destroy it     */
} vx_status;

/* BlockId (core/block_id.rs:18-30): leaf<<63 | types<<55 | mask<<47 | generation<<32 | index */
typedef uint64_t vx_block_id;
#define VX_BLOCK_EMPTY ((vx_block_id)0)             /* block_id.rs:112 */
#define VX_BLOCK_INVALID (~(vx_block_id)0)          /* block_id.rs:96  */

typedef struct vx_interner vx_interner; /* VoxInterner<T>   interner/mod.rs:25-40      */
typedef struct vx_tree vx_tree;         /* VoxTree<T>       spatial/voxtree.rs:108-113 */
typedef struct vx_batch vx_batch;       /* Batch<T>         core/batch.rs:39-45        */

/* InternerStats (interner/stats.rs:1-28), same field order. */
typedef struct vx_stats {
    uint64_t requested_budget, actual_budget, node_size, nodes_capacity;
    uint64_t total_allocations, total_deallocations, allocated_nodes, recycled_nodes;
    uint64_t alive_nodes, patterns;
    uint64_t total_cache_hits, total_cache_misses;
    uint64_t branch_cache_hits, branch_cache_misses, leaf_cache_hits, leaf_cache_misses;
    uint64_t collapsed_branches, leaf_nodes, branch_nodes;
    uint64_t max_alive_nodes, max_node_id, max_branch_ref_count, max_leaf_ref_count;
    uint64_t max_generation, generations_overflows;
} vx_stats;

/* Device memory an interner really holds (the reference's `actual_budget` only counts the node pools): pools =
 * capacity x (78 + sizeof T + 4 B free-list entry), the hash tables on top of the budget, and the scratch the builders
 * have grown to (level lists of the bulk builder, staging slabs of host batches, release frontiers). */
typedef struct vx_memory {
    uint64_t pools_bytes, table_bytes, leaf_table_bytes, bulk_scratch_bytes, stage_bytes, other_scratch_bytes;
    uint64_t pinned_host_bytes, total_device_bytes;
} vx_memory;

const char* vx_last_error(void);
int vx_abi_version(void);
int vx_device_count(void); /* number of CUDA devices, <0 on error */

/* ---------------------------------------------------------------- VoxInterner ------------- */
/* VoxInterner::<T>::with_memory_budget(bytes) — interner/mod.rs:45-155.
 * capacity = budget / node_size, node_size = 78 + sizeof(T) (mod.rs:158-164); index 0 is the
 * permanent empty branch.  All pools live in device memory of `device`; the hash table
 * (16 B per node of capacity) is allocated on top of the budget.  NULL on failure. */
vx_interner* vx_interner_create(size_t budget_bytes, vx_dtype dtype, int device);
void vx_interner_destroy(vx_interner*);
/* Back to the freshly-created state (every node dropped).  Trees built in it become dangling. */
int vx_interner_reset(vx_interner*);
/* Same, queued on `stream` (a cudaStream_t; NULL = the interner's stream) without synchronising. */
int vx_interner_reset_async(vx_interner*, void* stream);
size_t vx_interner_capacity(const vx_interner*);
vx_dtype vx_interner_dtype(const vx_interner*);
int vx_interner_device(const vx_interner*);
/* Number of index slots handed out so far, including slot 0 (`next_index`, mod.rs:28). */
int64_t vx_interner_next_index(const vx_interner*);
/* VoxInterner::get_ref — mod.rs:218-228 */
int vx_interner_get_ref(const vx_interner*, vx_block_id id, uint32_t* out);
/* VoxInterner::get_value / get_children — mod.rs:166-203 */
int vx_interner_get_value(const vx_interner*, vx_block_id id, int64_t* out);
int vx_interner_get_children(const vx_interner*, vx_block_id id, vx_block_id out[8]);
/* InternerStats (feature memory_stats) */
int vx_interner_stats(const vx_interner*, vx_stats* out);
int vx_interner_memory(const vx_interner*, vx_memory* out);
/* Diagnostic: the bulk builder's unit memo — out[0] entries inserted since the last wipe, out[1] occupied slots,
 * out[2], out[3] the first occupied entry. */
int vx_interner_debug_memo(const vx_interner*, uint64_t out[4]);
/* Copies the first `next_index` entries of each pool to HOST arrays (any may be NULL):
 * children[n][8], values[n] (widened to int64), ref_counts[n], generations[n], hashes[n].
 * `cap` = entries available in the caller's arrays; returns next_index or <0. */
int64_t vx_interner_download(const vx_interner*, size_t cap, vx_block_id* children, int64_t* values,
                             uint32_t* ref_counts, uint16_t* generations, uint64_t* hashes);
/* Blocks until all work queued on the interner's stream has finished; surfaces deferred errors. */
int vx_interner_sync(vx_interner*);
/* The interner's CUDA stream (a cudaStream_t), for callers that want to order their own work. */
void* vx_interner_stream(const vx_interner*);

/* ---------------------------------------------------------------- Batch ------------------- */
/* Batch::new(max_depth) / VoxTree::create_batch — core/batch.rs:63-81, voxtree.rs:296-301.
 * Host arrays (pinned): masks[B][2] = (set_mask, clear_mask), values[B][8], B = 8^(max_depth-1),
 * block p = Morton(x>>1,y>>1,z>>1), lane i = (x&1)|(y&1)<<1|(z&1)<<2 (utils/common.rs:24-55). */
vx_batch* vx_batch_create(uint8_t max_depth, vx_dtype dtype);
void vx_batch_destroy(vx_batch*);
/* Batch::set / just_set — batch.rs:145-175,211-213.  Returns 1 (state changed), VX_E_BOUNDS when
 * the position is outside the chunk (reference: debug_assert). */
int vx_batch_set(vx_batch*, int x, int y, int z, int64_t voxel);
/* Batch::fill / just_fill — batch.rs:178-184,218-221 (resets recorded patches). */
int vx_batch_fill(vx_batch*, int64_t value);
/* Batch::clear / just_clear — batch.rs:187-195,223-225 */
int vx_batch_clear(vx_batch*);
/* Batch::masks / values / to_fill / size / has_patches — batch.rs:86-132.  The raw arrays may be
 * bulk-written by the caller; call vx_batch_mark_patched afterwards.  A set bit whose value is the
 * default (0) cannot be produced by Batch::set and is ignored by apply. */
uint8_t* vx_batch_masks(vx_batch*);
void* vx_batch_values(vx_batch*);
size_t vx_batch_blocks(const vx_batch*);
int vx_batch_to_fill(const vx_batch*, int64_t* out); /* 1 = Some(*out), 0 = None */
size_t vx_batch_size(const vx_batch*);
int vx_batch_has_patches(const vx_batch*);
void vx_batch_mark_patched(vx_batch*);
/* Array form of Batch::set (one FFI crossing for n voxels): xyz[n][3], voxels[n]; stops at the first error. */
int vx_batch_set_many(vx_batch*, size_t n, const int32_t* xyz, const int64_t* voxels);
/* Replaces the batch's content with dense arrays in Batch layout (masks[B][2], values[B][8]) — what a
 * caller holding a reference `Batch` hands over (`batch.masks()`, `batch.values()`, batch.rs:86-105);
 * to_fill becomes None, has_patches = any mask bit. */
int vx_batch_assign(vx_batch*, const uint8_t* masks, const void* values);
/* Next to has_patches (batch.rs:44) a batch records which units of 512 Morton-consecutive blocks
 * Batch::set ever touched; vx_apply_batches moves only those across the bus.  Returns their number. */
int vx_batch_touched_units(const vx_batch*);
uint8_t vx_batch_max_depth(const vx_batch*);
vx_dtype vx_batch_dtype(const vx_batch*);

/* ---------------------------------------------------------------- VoxTree ----------------- */
/* VoxTree::new(MaxDepth) — voxtree.rs:116-126.  max_depth in [2,7]; the reference asserts < 7
 * (core/max_depth.rs:77-83), depth 7 is this build's extension.  NULL on failure. */
vx_tree* vx_tree_create(uint8_t max_depth);
void vx_tree_destroy(vx_tree*);
vx_block_id vx_tree_root_id(const vx_tree*);          /* get_root_id  voxtree.rs:128-133 */
/* set_root_id (voxtree.rs:135-141): adopts `root` and bumps its refcount. */
int vx_tree_set_root_id(vx_interner*, vx_tree*, vx_block_id root);
uint8_t vx_tree_max_depth(const vx_tree*);            /* VoxOpsConfig  voxtree.rs:331-341 */
uint32_t vx_tree_voxels_per_axis(const vx_tree*);
int vx_tree_is_empty(const vx_tree*);                 /* VoxOpsState   voxtree.rs:343-353 */
int vx_tree_is_leaf(const vx_tree*);
int vx_tree_is_dirty(const vx_tree*);                 /* VoxOpsDirty   voxtree.rs:355-370 */
void vx_tree_mark_dirty(vx_tree*);
void vx_tree_clear_dirty(vx_tree*);
/* Hands the tree a root whose reference the caller already holds (the roots vx_model_deserialize returns carry
 * the reference VoxTree::set_root_id took for them, voxtree.rs:135-141).  No refcount change; marks the tree dirty.
 * The id is NOT checked (there is no interner in this call): use vx_tree_set_root_id for ids of unknown provenance —
 * it verifies that the slot is live and of the id's generation (is_valid_block_id, interner/mod.rs:997-1008). */
int vx_tree_adopt_root(vx_tree*, vx_block_id root);
/* After vx_interner_reset every tree built in that interner dangles; this makes n of them empty again
 * without touching the interner (what dropping and re-creating the VoxTrees does in the reference). */
int vx_trees_forget(vx_tree* const* trees, size_t n);

/* VoxTree::apply_batch — voxtree.rs:303-328 (set_batch_at_root :708-722,
 * set_batch_at_depth_iterative :724-1118).  Returns 1 = changed, 0 = unchanged, <0 error.
 * Synchronous: on return tree->root is final. */
int vx_tree_apply_batch(vx_interner*, vx_tree*, const vx_batch*);

/* NEW multi-chunk entry (replaces the serial loop voxelis-voxelize/src/lib.rs:357-361):
 * result == applying batches[i] to trees[i] in index order into one interner, up to a
 * permutation of node indices.  `changed` (may be NULL) receives 1/0 per tree. */
int vx_apply_batches(vx_interner*, vx_tree* const* trees, const vx_batch* const* batches, size_t n,
                     uint8_t* changed);
/* Host batches are read in place (pinned + mapped arena slots): only the touched units' masks and the
 * values of blocks with a set bit cross PCIe (vx_stage.cuh), so a sparse world moves a few percent of its
 * batch bytes.  The device slab is bounded by VX_STAGE_MAX_BYTES (default 4 GiB; larger calls are sliced). */

/* Slab form of the same for FRESH trees: n chunks of depth `max_depth`, arrays laid out
 * masks[n][B][2], values[n][B][8].  `flags` (may be NULL = patches, no fill): bit0 = to_fill is
 * Some(fills[i]), bit1 = has_patches.  Pointers may be host (pinned or pageable) or device
 * memory of the interner's device — detected per pointer.  roots_out[n] / changed_out[n] (may be
 * NULL) likewise.  Host inputs are streamed through double-buffered device staging. */
#define VX_FLAG_FILL 1u
#define VX_FLAG_PATCHES 2u
int vx_apply_batches_slab(vx_interner*, uint8_t max_depth, size_t n, const uint8_t* masks,
                          const void* values, const uint8_t* flags, const int64_t* fills,
                          vx_block_id* roots_out, uint8_t* changed_out);
/* Fully asynchronous device-resident form: every pointer is device memory, work is queued on
 * `stream` (NULL = the interner's stream) and nothing is synchronised.  Errors surface at the
 * next vx_interner_sync(). */
int vx_apply_batches_device(vx_interner*, uint8_t max_depth, size_t n, const uint8_t* d_masks,
                            const void* d_values, const uint8_t* d_flags, const int64_t* d_fills,
                            vx_block_id* d_roots_out, uint8_t* d_changed_out, void* stream);

/* Batch generation in device memory (the step before the path): generate_terrain_batch /
 * generate_terrain_batch_3_mats — utils/shapes.rs:273-357 for a whole grid of chunks, writing the Batch arrays of
 * vx_apply_batches_device directly in HBM.  Both calls are asynchronous on `stream` (NULL = the interner's stream).
 *   vx_terrain_heights_device  d_heights[nx][nz] (int32) in [0, height): this library's seeded integer value noise
 *                              (the reference samples fastnoise-lite OpenSimplex2, a float third-party noise that is
 *                              input synthesis, not part of the path); column (x0 + i, z0 + j)
 *   vx_terrain_batches_device  grid = chunks along (x, y, z), chunk index (cx*gy + cy)*gz + cz, d_heights sized
 *                              [gx*N][gz*N]; surface_only: the voxel at Y == h (shapes.rs:302-304), else every voxel
 *                              with Y <= h (:306-309); materials 3: 1 for Y >= h-2, 2 for Y >= h-4, 3 deeper
 *                              (:344-350); masks[n][B][2], values[n][B][8] of the interner's dtype */
int vx_terrain_heights_device(vx_interner*, uint32_t nx, uint32_t nz, uint64_t seed, uint32_t height, int64_t x0,
                              int64_t z0, int32_t* d_heights, void* stream);
int vx_terrain_batches_device(vx_interner*, uint8_t max_depth, const uint32_t grid[3], const int32_t* d_heights,
                              int surface_only, int materials, uint8_t* d_masks, void* d_values, void* stream);
/* High-entropy benchmark batches written in device memory (BASELINE config 2 "high-entropy" / config 4; the reference
 * fills such batches voxel by voxel from a PRNG before timing, voxtree_bench.rs:553-563 pattern): n chunks,
 * v = 1 + splitmix64(cell_index + ((seed_base + chunk0 + i) << 32)) mod k per cell of cell^3 voxels; k == 4 gives
 * splitmix64 mod 4 with 0 recorded as a clear.  Asynchronous on `stream`. */
int vx_random_batches_device(vx_interner*, uint8_t max_depth, size_t n, uint64_t seed_base, uint64_t chunk0, uint32_t k,
                             uint32_t cell, uint8_t* d_masks, void* d_values, void* stream);

/* The voxeliser's two steps (voxelis-voxelize/src/lib.rs), SURVEY §8f-4:
 *   vx_voxelize_plan           Voxelizer::build_face_to_chunk_map (:113-156), host work as in the reference: which faces
 *                              (faces[nf][3], 1-based vertex indices, vertices[nv][3] f64) touch which chunk.  Returns
 *                              the number of chunks and *n_pairs_out; positions_out[n][3] (chunks in first-seen order —
 *                              the reference keeps a hash map) and the flat (chunk, face) pair lists are filled when
 *                              the capacities suffice (call once with NULL outputs for the sizes).
 *   vx_voxelize_chunks_device  Voxelizer::voxelize_chunk (:159-249) with triangle_cube_intersection
 *                              (voxelis-math/src/lib.rs:3-214, f64, the reference's order of operations) for ALL planned
 *                              chunks: zeroes and fills the device slab masks[n][B][2] / values[n][B][8] (value 1, of
 *                              the interner's dtype) that vx_apply_batches_device reads; d_has_patches[n] (device, may
 *                              be NULL) = Batch::has_patches — the reference drops chunks without patches (:245-249).
 *                              Host arrays are copied in; the call synchronises. */
int64_t vx_voxelize_plan(uint8_t max_depth, double chunk_world_size, const double mesh_min[3], size_t n_vertices,
                         const double* vertices, size_t n_faces, const int32_t* faces, int32_t* positions_out,
                         size_t cap_chunks, uint32_t* pair_chunk_out, uint32_t* pair_face_out, size_t cap_pairs,
                         size_t* n_pairs_out);
int vx_voxelize_chunks_device(vx_interner*, uint8_t max_depth, double chunk_world_size, const double mesh_min[3],
                              size_t n_vertices, const double* vertices, size_t n_faces, const int32_t* faces,
                              size_t n_chunks, const int32_t* positions, size_t n_pairs, const uint32_t* pair_chunk,
                              const uint32_t* pair_face, uint8_t* d_masks, void* d_values, uint8_t* d_has_patches);

/* VoxTree::get — voxtree.rs:144-160 -> get_at_depth utils/common.rs:122-156.
 * Returns 1 = Some(*out), 0 = None, VX_E_BOUNDS outside the chunk (reference: assert!). */
int vx_tree_get(const vx_interner*, const vx_tree*, int x, int y, int z, int64_t* out);
/* Bulk get: xyz[n][3] (host), found[n], values[n] (host). */
int vx_tree_get_many(const vx_interner*, const vx_tree*, size_t n, const int32_t* xyz, uint8_t* found,
                     int64_t* values);
/* to_vec — utils/common.rs:158-246: dense T[N^3], index = y*N*N + z*N + x.  `dense` host or device. */
int vx_tree_to_vec(const vx_interner*, const vx_tree*, void* dense);
/* to_vec for n bare roots of one depth: dense[n][N^3]; roots host or device, dense host or device. */
int vx_roots_to_vec(const vx_interner*, uint8_t max_depth, size_t n, const vx_block_id* roots, void* dense);
/* to_vec at a level of detail — to_vec(interner, root, max_depth.for_lod(lod)), world/voxchunk.rs:267 with
 * core/max_depth.rs:137-140: only max_depth - lod levels (saturating) are unfolded and a branch standing at
 * that depth contributes its LOD value (calc_average, core/voxel.rs:96-141).  dense[n][M^3], M = 2^(max_depth-lod). */
int vx_roots_to_vec_lod(const vx_interner*, uint8_t max_depth, uint8_t lod, size_t n, const vx_block_id* roots,
                        void* dense);

/* generate_occupancy_masks — utils/mesh.rs:515-596 (+ fill_masks_for_region :418-513 and the ordering of
 * OccupancyDataBuilder::build :263-285): the greedy mesher's bit planes, computed from the device pools.
 * n chunks (roots[n], host) are unfolded to max_depth - lod levels (saturating, <= 6: a builder covers 64^3 voxels,
 * mesh.rs:50) and placed at offsets[n][3] (host; voxels at that level of detail, multiples of the chunk side — the
 * `offset` argument of the reference) of builder builder_of[n] (host, NULL = all in builder 0); each call of the
 * reference with the same builder = one entry with the same builder index.  Outputs, all host or all device:
 *   global[nb][3*4096]           YZ word[y*64+z] bit x | XZ 4096 + word[z*64+x] bit y | XY 8192 + word[y*64+x] bit z
 *   active[nb][6]                global_active (mesh.rs:451-461)
 *   n_materials[nb], material_ids[nb][max_materials], material_counts[nb][max_materials]
 *                                `materials` sorted by id (id = value as usize, core/voxel.rs:85-87), voxel counts
 *   per_material[nb][max_materials][3*4096]   only the first n_materials[b] planes of builder b are written
 * VX_E_BOUNDS when a builder holds more than max_materials (<= 1024) materials; n_materials[] is valid then.
 * At most 65535 builders per call.  Builders with <= min(max_materials, 5) materials and chunks of at least 8^3 voxels
 * take the shared-memory kernel (pass the smallest max_materials that fits: <= 3 keeps two CTAs per SM). */
int vx_occupancy_masks(const vx_interner*, uint8_t max_depth, uint8_t lod, size_t n, const vx_block_id* roots,
                       const uint32_t* offsets, const uint32_t* builder_of, size_t n_builders, uint32_t max_materials,
                       uint64_t* global, uint64_t* active, uint32_t* n_materials, uint64_t* material_ids,
                       uint64_t* material_counts, uint64_t* per_material);

/* VoxTree::fill / clear — voxtree.rs:264-292. */
int vx_tree_fill(vx_interner*, vx_tree*, int64_t value);
int vx_tree_clear(vx_interner*, vx_tree*);

/* ---------------------------------------------------------------- VTM export -------------- */
/* VoxModel::serialize — world/voxmodel.rs:177-294 (+ serialize_chunk, world/voxchunk.rs:382-405): the VTM payload
 * of n chunks (positions[n][3] chunk coordinates, roots[n], both host) sharing this interner, written from the
 * device pools: every alive node renumbered leaves-first in index order, records as the reference writes them,
 * then the chunk table in the order given (the reference walks a hash map).  Returns the payload size (also when
 * `out` is NULL or `cap` is too small, in which case nothing is written), <0 on error. */
int64_t vx_model_serialize(const vx_interner*, size_t n, const int32_t* positions, const vx_block_id* roots,
                           uint8_t* out, size_t cap);
/* export_model_to_vtm — io/export.rs:90-151: header, MD5 of the payload, payload (zstd level 7 when `compress`,
 * Flags::COMPRESSED; libzstd is looked up at run time, VX_E_UNSUPPORTED if absent) -> `path`.  The file opens
 * with the reference's import_model_from_vtm (io/import.rs:14-98). */
int vx_export_vtm(const vx_interner*, const char* path, const char* name, uint8_t max_depth, float chunk_world_size,
                  const int32_t world_bounds[3], size_t n, const int32_t* positions, const vx_block_id* roots,
                  int compress);

/* VoxModel::deserialize — world/voxmodel.rs:296-408 (+ deserialize_chunk, world/voxchunk.rs:407-440; interner side
 * interner/mod.rs:930-1000) into a FRESH interner: node k of the file becomes pool index k, as the reference asserts.
 * Returns the number of chunks and writes positions_out[n][3] / roots_out[n] (host; each root holds one reference,
 * voxtree.rs:135-141); fails with VX_E_INVALID before touching the interner when `cap` < n or the data is malformed. */
int64_t vx_model_deserialize(vx_interner*, const uint8_t* data, size_t len, int32_t* positions_out,
                             vx_block_id* roots_out, size_t cap);
typedef struct vx_vtm_info {
    uint16_t flags;          /* io/flags.rs: bit 0 = COMPRESSED */
    uint8_t max_depth;       /* "lod_level" in io/import.rs:36 */
    float chunk_world_size;
    int32_t world_bounds[3];
    char name[256];
} vx_vtm_info;
/* import_model_from_vtm — io/import.rs:14-98: header, zstd (when flagged), MD5 check, then vx_model_deserialize. */
int64_t vx_import_vtm(vx_interner*, const char* path, vx_vtm_info* info_out, int32_t* positions_out,
                      vx_block_id* roots_out, size_t cap);

/* ---------------------------------------------------------------- multi-GPU world (new) ---- */
/* One process per GPU.  The reference keeps ONE interner for the whole world (world/voxmodel.rs:31-32); here every
 * GPU builds its chunks into a private interner with no communication (vx_apply_batches*), and
 * vx_world_global_dedup optionally merges the private interners into hash-partitioned GLOBAL shards — the Voxelis
 * Bible's "shard the pattern map by hash % N" (§3.9, §13): height-synchronous rounds, records travel to
 * owner = hash(children as global ids) mod G over NCCL send / recv (NVLink), owners intern them
 * (get_or_create_branch / _leaf, interner/mod.rs:627-829) and answer with 8-byte global ids.
 *   global id = the owner shard's BlockId with the owner's rank in generation bits [44..46].
 * NCCL is loaded at run time (libnccl.so.2); without it these entries return VX_E_UNSUPPORTED.
 *   vx_world_unique_id     rank 0 makes the 128-byte rendezvous id; the host passes it to the other processes
 *   vx_world_create        collective over the n_ranks processes (ncclCommInitRank), n_ranks <= 8
 *   vx_world_barrier       device-side barrier across the ranks on `stream`, no host synchronisation
 *   vx_world_global_dedup  collective: `local` = this rank's private interner, `shard` = this rank's (empty) global
 *                          shard, roots[n_roots] = this rank's chunk roots (host) -> global_roots_out (host).
 *                          Sum over ranks of the shard sizes == the node count of ONE shared interner. */
#define VX_WORLD_ID_BYTES 128
typedef struct vx_world vx_world;
typedef struct vx_dedup_summary {
    uint64_t rounds, branches, leaves;          /* global unique nodes over all shards */
    uint64_t bytes_sent;                        /* all ranks, records + ids */
    uint64_t local_nodes_all_ranks;             /* sum of the private interners' node counts (before the merge) */
    uint64_t this_shard_branches, this_shard_leaves;
} vx_dedup_summary;
int vx_world_unique_id(uint8_t out[VX_WORLD_ID_BYTES]);
vx_world* vx_world_create(int n_ranks, int rank, const uint8_t id[VX_WORLD_ID_BYTES], int device);
void vx_world_destroy(vx_world*);
int vx_world_size(const vx_world*);
int vx_world_rank(const vx_world*);
int vx_world_barrier(vx_world*, void* stream);
int vx_world_global_dedup(vx_world*, vx_interner* local, vx_interner* shard, size_t n_roots, const vx_block_id* roots,
                          vx_block_id* global_roots_out, vx_dedup_summary* out);

/* ---------------------------------------------------------------- global dedup (new) ------ */
/* Optional merge of per-GPU interners into hash-partitioned global shards (the Voxelis Bible's
 * "shard the pattern map by hash % N", §3.9/§13).  Height-synchronous rounds: pack the local
 * nodes of one height by owner = hash(children as global ids) mod G, exchange (NCCL all-to-all, done
 * by the caller), intern on the owner, send the global ids back.  global id = the owner shard's BlockId
 * with the owner's rank in generation bits [44..46].  All pointers are DEVICE memory; calls synchronise.
 *   vx_dedup_heights            d_heights[next_index]: 0 leaf, h>=1 branch, 255 unused; returns max height
 *   vx_dedup_pack               counts_out[G] (host); with d_records != NULL also writes the 72-byte records
 *                               (8 global child ids + value) grouped by owner and d_src[j] = local node index
 *   vx_dedup_scatter            d_gmap[d_src[j]] = d_ids[j]
 *   vx_dedup_map_roots          d_out[j] = d_gmap[index(d_roots[j])]
 *   vx_interner_intern_records  owner side: get_or_create for n records; *created_out = new nodes */
int vx_dedup_heights(vx_interner*, uint8_t* d_heights);
int vx_dedup_pack(vx_interner*, int height, const uint8_t* d_heights, const uint64_t* d_gmap, int G,
                  uint64_t* counts_out, uint64_t* d_records, uint32_t* d_src);
int vx_dedup_scatter(vx_interner*, size_t n, const uint32_t* d_src, const uint64_t* d_ids, uint64_t* d_gmap);
int vx_dedup_map_roots(vx_interner*, size_t n, const uint64_t* d_roots, const uint64_t* d_gmap, uint64_t* d_out);
int vx_interner_intern_records(vx_interner* shard, size_t n, const uint64_t* d_records, int rank, int leaf_round,
                               uint64_t* d_ids_out, uint64_t* created_out);

#ifdef __cplusplus
}
#endif
#endif /* VOXELIS_B200_H */

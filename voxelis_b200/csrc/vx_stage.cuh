// vx_stage.cuh — gather of SPARSE host batches into the device slab the builders read.
//
// A vx_batch lives in pinned, mapped host memory (one slot of the batch arena, vx_capi.cu) and keeps,
// next to the reference's `has_patches` flag (core/batch.rs:44), one bit per 512-block unit saying
// whether Batch::set ever touched it.  vx_apply_batches turns those bits into a flat list of touched
// units; this kernel pulls exactly those units across PCIe ("zero copy" loads issued by the SMs):
//   * the unit's masks, 1 KiB, one 32-byte load per lane;
//   * the values of the blocks that have a set bit (phase 1 never looks at the others,
//     spatial/voxtree.rs:779-781), one 8-byte (u8) / 32-byte (i32) load per block, lane = block so
//     neighbouring blocks coalesce into one request.
// Untouched units are never read: their masks in the slab were zeroed by a memset in HBM.  A sparse
// world therefore moves a few percent of its batch bytes over the bus (perlin surface world: ~50 MB of
// 1.34 GB) and the copy engine is not involved at all.
//
// stage_units_occ_kernel goes one step further for batches that were only ever written through the
// Batch API (set / fill / clear / assign): there `value != 0  <=>  set bit` (batch.rs:162-168 writes both
// together), so the set_mask need not travel at all.  The batch keeps ONE BIT PER BLOCK ("some voxel of
// this block was set"), 64 B per unit instead of 1 KiB of masks; the kernel pulls that bitmap, then the
// values of the flagged blocks, and rebuilds each block's set_mask from its values in registers.
#pragma once
#include "vx_device.cuh"

namespace vx {

constexpr int STAGE_THREADS = 256;
constexpr int STAGE_WARPS = STAGE_THREADS / 32;

// VB = bytes of one block's eight values (8 for u8, 32 for i32).
template <int VB>
__global__ void __launch_bounds__(STAGE_THREADS)
stage_units_kernel(const u64* __restrict__ src,    // [n] device-visible address of each batch slot (masks, then values)
                   const u32* __restrict__ units,  // [n_units] chunk * units_per_chunk + unit
                   u32 n_units, u32 upc_log2, u32 unit_blocks, u8* __restrict__ d_masks, u8* __restrict__ d_values,
                   u64 mask_bytes, u64 value_bytes) {
    __shared__ __align__(16) u8 sm_masks[STAGE_WARPS][1024];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 warps = gridDim.x * STAGE_WARPS;
    u8* my = sm_masks[w];
    for (u32 e = blockIdx.x * STAGE_WARPS + w; e < n_units; e += warps) {
        const u32 u = units[e];
        const u32 chunk = u >> upc_log2, unit = u & ((1u << upc_log2) - 1u);
        const u8* base = reinterpret_cast<const u8*>(src[chunk]);
        const u64 first = u64(unit) * unit_blocks;  // first block of the unit inside its chunk
        const bool mine = u32(lane) * 16u < unit_blocks;
        uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
        if (mine) {
            const u8* p = base + first * 2 + lane * 32;
            m0 = ld_stream_v4(p);
            if (unit_blocks >= 16) m1 = ld_stream_v4(p + 16);  // D = 2 has only 8 blocks (16 bytes of masks)
        }
        reinterpret_cast<uint4*>(my)[lane * 2] = m0;
        reinterpret_cast<uint4*>(my)[lane * 2 + 1] = m1;
        if (mine) {
            u8* q = d_masks + u64(chunk) * mask_bytes + first * 2 + lane * 32;
            *reinterpret_cast<uint4*>(q) = m0;
            if (unit_blocks >= 16) *reinterpret_cast<uint4*>(q + 16) = m1;
        }
        __syncwarp();
        const u8* vs = base + mask_bytes + first * VB;
        u8* vd = d_values + u64(chunk) * value_bytes + first * VB;
        const u32 iters = (unit_blocks + 31) >> 5;
        if (VB == 8) {
            // all (<= 16) iterations of the unit in flight at once: the bus round trip is microseconds
            uint2 r[16];
            u32 have = 0;
#pragma unroll
            for (u32 it = 0; it < 16; ++it) {
                const u32 b = it * 32 + lane;
                if (it < iters && b < unit_blocks && my[b * 2] != 0) {
                    const u64 v = ld_stream_u64(vs + u64(b) * 8);
                    r[it] = make_uint2(u32(v), u32(v >> 32));
                    have |= 1u << it;
                }
            }
#pragma unroll
            for (u32 it = 0; it < 16; ++it)
                if (have >> it & 1) *reinterpret_cast<uint2*>(vd + u64(it * 32 + lane) * 8) = r[it];
        } else {
            for (u32 g = 0; g < iters; g += 4) {
                uint4 r[8];
                u32 have = 0;
#pragma unroll
                for (u32 k = 0; k < 4; ++k) {
                    const u32 b = (g + k) * 32 + lane;
                    if (g + k < iters && b < unit_blocks && my[b * 2] != 0) {
                        r[2 * k] = ld_stream_v4(vs + u64(b) * 32);
                        r[2 * k + 1] = ld_stream_v4(vs + u64(b) * 32 + 16);
                        have |= 1u << k;
                    }
                }
#pragma unroll
                for (u32 k = 0; k < 4; ++k)
                    if (have >> k & 1) {
                        uint4* q = reinterpret_cast<uint4*>(vd + u64((g + k) * 32 + lane) * 32);
                        q[0] = r[2 * k];
                        q[1] = r[2 * k + 1];
                    }
            }
        }
        __syncwarp();  // the next unit overwrites this warp's mask tile
    }
}


__device__ __forceinline__ u32 nonzero_bytes(u64 v) {  // bit i = byte i of v is not 0
    v |= v >> 4;
    v |= v >> 2;
    v |= v >> 1;
    v &= 0x0101010101010101ull;
    return u32((v * 0x0102040810204080ull) >> 56);
}

template <int VB>
__global__ void __launch_bounds__(STAGE_THREADS)
stage_units_occ_kernel(const u64* __restrict__ src,    // [n] slot address: masks, values, then the block bitmap
                       const u32* __restrict__ units, u32 n_units, u32 upc_log2, u32 unit_blocks,
                       u8* __restrict__ d_masks, u8* __restrict__ d_values, u64 mask_bytes, u64 value_bytes) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 warps = gridDim.x * STAGE_WARPS;
    const u32 iters = (unit_blocks + 31) >> 5;
    for (u32 e = blockIdx.x * STAGE_WARPS + w; e < n_units; e += warps) {
        const u32 u = units[e];
        const u32 chunk = u >> upc_log2, unit = u & ((1u << upc_log2) - 1u);
        const u8* base = reinterpret_cast<const u8*>(src[chunk]);
        const u64 first = u64(unit) * unit_blocks;
        u32 word = 0;  // lane L < iters: the bits of blocks [32 L, 32 L + 32) of this unit
        if (u32(lane) < iters) {
            const u8* occ = base + mask_bytes + value_bytes + (first >> 3);  // D = 2: one byte, the rest is padding
            asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(word) : "l"(occ + 4 * lane));
        }
        const u8* vs = base + mask_bytes + first * VB;
        u8* vd = d_values + u64(chunk) * value_bytes + first * VB;
        u16* md = reinterpret_cast<u16*>(d_masks + u64(chunk) * mask_bytes + first * 2);
        if (VB == 8) {
            uint2 r[16];
            u32 have = 0;
#pragma unroll
            for (u32 it = 0; it < 16; ++it) {
                const u32 bits = __shfl_sync(0xFFFFFFFFu, word, it);
                const u32 b = it * 32 + lane;
                if (it < iters && b < unit_blocks && (bits >> lane & 1)) {
                    const u64 v = ld_stream_u64(vs + u64(b) * 8);
                    r[it] = make_uint2(u32(v), u32(v >> 32));
                    have |= 1u << it;
                }
            }
#pragma unroll
            for (u32 it = 0; it < 16; ++it) {
                const u32 b = it * 32 + lane;
                u32 m = 0;
                if (have >> it & 1) {
                    *reinterpret_cast<uint2*>(vd + u64(b) * 8) = r[it];
                    m = nonzero_bytes(u64(r[it].x) | u64(r[it].y) << 32);
                }
                // every block of a staged unit gets its masks (set_mask, clear_mask = 0): the builders may
                // then read the unit without the slab having been zeroed first
                if (it < iters && b < unit_blocks) md[b] = u16(m);
            }
        } else {
            for (u32 g = 0; g < iters; g += 4) {
                uint4 r[8];
                u32 have = 0;
#pragma unroll
                for (u32 k = 0; k < 4; ++k) {
                    const u32 bits = __shfl_sync(0xFFFFFFFFu, word, (g + k) & 31);
                    const u32 b = (g + k) * 32 + lane;
                    if (g + k < iters && b < unit_blocks && (bits >> lane & 1)) {
                        r[2 * k] = ld_stream_v4(vs + u64(b) * 32);
                        r[2 * k + 1] = ld_stream_v4(vs + u64(b) * 32 + 16);
                        have |= 1u << k;
                    }
                }
#pragma unroll
                for (u32 k = 0; k < 4; ++k) {
                    const u32 b = (g + k) * 32 + lane;
                    u32 m = 0;
                    if (have >> k & 1) {
                        uint4* q = reinterpret_cast<uint4*>(vd + u64(b) * 32);
                        const uint4 a = r[2 * k], c = r[2 * k + 1];
                        q[0] = a;
                        q[1] = c;
                        m = (a.x != 0) | (a.y != 0) << 1 | (a.z != 0) << 2 | (a.w != 0) << 3 | (c.x != 0) << 4 |
                            (c.y != 0) << 5 | (c.z != 0) << 6 | (c.w != 0) << 7;
                    }
                    if (g + k < iters && b < unit_blocks) md[b] = u16(m);
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Journal path (vx_batch: jblock / jvals).  A batch that was only written through the API keeps the blocks holding a
// non-default voxel PACKED: 2 bytes of block index + VB bytes of values per flagged block, contiguous in its pinned
// slot.  stage_journal_kernel streams those arrays over the bus with fully coalesced loads (32 lanes x 8 B = 256 B per
// request for u8) and scatters them into the slab; the set_mask is rebuilt from the values (value != 0 <=> set bit for
// such batches, batch.rs:162-168).  Against the occupancy path — 8-byte reads scattered over the unit, which the bus
// moves as 32-byte sectors — the perlin world sends 13.8 MB instead of ~30.  The masks of the OTHER blocks of a
// touched unit must read zero: stage_zero_units_kernel clears the listed units' masks in HBM first.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(STAGE_THREADS)
stage_zero_units_kernel(const u32* __restrict__ units, u32 n_units, u32 upc_log2, u32 unit_blocks, u8* __restrict__ d_masks,
                        u64 mask_bytes) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 warps = gridDim.x * STAGE_WARPS;
    const u32 bytes = unit_blocks * 2;  // a multiple of 16 (unit_blocks >= 64 on this path)
    for (u32 e = blockIdx.x * STAGE_WARPS + w; e < n_units; e += warps) {
        const u32 u = units[e];
        const u32 chunk = u >> upc_log2, unit = u & ((1u << upc_log2) - 1u);
        u8* q = d_masks + u64(chunk) * mask_bytes + u64(unit) * bytes;
        for (u32 off = u32(lane) * 16; off < bytes; off += 512) *reinterpret_cast<uint4*>(q + off) = make_uint4(0, 0, 0, 0);
    }
}

template <int VB>
__global__ void __launch_bounds__(STAGE_THREADS)
stage_journal_kernel(const u64* __restrict__ src,   // [n] device-visible address of each batch slot
                     const u32* __restrict__ jcnt,  // [n] journal entries of each batch
                     u32 n_chunks, u64 off_jvals, u64 off_jblock, u8* __restrict__ d_masks, u8* __restrict__ d_values,
                     u64 mask_bytes, u64 value_bytes) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 warps = gridDim.x * STAGE_WARPS;
    for (u32 chunk = blockIdx.x * STAGE_WARPS + w; chunk < n_chunks; chunk += warps) {
        const u32 cnt = jcnt[chunk];
        if (cnt == 0) continue;
        const u8* base = reinterpret_cast<const u8*>(src[chunk]);
        const u8* jv = base + off_jvals;
        const u8* jb = base + off_jblock;
        u8* vd = d_values + u64(chunk) * value_bytes;
        u16* md = reinterpret_cast<u16*>(d_masks + u64(chunk) * mask_bytes);
        constexpr u32 UNR = VB == 8 ? 8 : 4;  // entries per lane in flight: the bus round trip is microseconds
        for (u32 k0 = 0; k0 < cnt; k0 += 32 * UNR) {
            u32 blk[UNR];
            if (VB == 8) {
                u64 v[UNR];
#pragma unroll
                for (u32 j = 0; j < UNR; ++j) {
                    const u32 k = k0 + j * 32 + lane;
                    blk[j] = 0xFFFFFFFFu;
                    v[j] = 0;
                    if (k < cnt) {
                        blk[j] = ld_stream_u16(jb + 2 * u64(k));
                        v[j] = ld_stream_u64(jv + 8 * u64(k));
                    }
                }
#pragma unroll
                for (u32 j = 0; j < UNR; ++j)
                    if (blk[j] != 0xFFFFFFFFu) {
                        *reinterpret_cast<u64*>(vd + u64(blk[j]) * 8) = v[j];
                        md[blk[j]] = u16(nonzero_bytes(v[j]));
                    }
            } else {
                uint4 a[UNR], c[UNR];
#pragma unroll
                for (u32 j = 0; j < UNR; ++j) {
                    const u32 k = k0 + j * 32 + lane;
                    blk[j] = 0xFFFFFFFFu;
                    if (k < cnt) {
                        blk[j] = ld_stream_u16(jb + 2 * u64(k));
                        a[j] = ld_stream_v4(jv + 32 * u64(k));
                        c[j] = ld_stream_v4(jv + 32 * u64(k) + 16);
                    }
                }
#pragma unroll
                for (u32 j = 0; j < UNR; ++j)
                    if (blk[j] != 0xFFFFFFFFu) {
                        uint4* q = reinterpret_cast<uint4*>(vd + u64(blk[j]) * 32);
                        q[0] = a[j];
                        q[1] = c[j];
                        md[blk[j]] = u16((a[j].x != 0) | (a[j].y != 0) << 1 | (a[j].z != 0) << 2 | (a[j].w != 0) << 3 |
                                         (c[j].x != 0) << 4 | (c[j].y != 0) << 5 | (c[j].z != 0) << 6 | (c[j].w != 0) << 7);
                    }
            }
        }
    }
}

}  // namespace vx

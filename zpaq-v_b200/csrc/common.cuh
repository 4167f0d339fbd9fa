// common.cuh -- device-side structures and helpers shared by the libzpaqgpu kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "model.h"

namespace zg {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

// Constant lookup tables in HBM (copied to shared memory by the chain kernel).
struct DevTables {
    const int16_t *stretch;  // 32768 entries, stretch_table narrowed to i16 (range +-2047)
    const u16 *squash;       // 4096 entries, squash_table narrowed to u16 (range 1..32767)
    const u16 *squash_pad;   // 4096 entries indexed by p+2048 for p in [-2048,2047]: the index clamp of
                             // squash() (predictor.v:193-202) folded into the table
    const int16_t *stretch_pad;  // stretch with entry 0 := entry 1 (stretch() clamps its index to >= 1)
    const u8 *nex;           // 512 entries: nex[s*2+y] = next state (statetable.v ns[s*4+y])
    const i32 *dt;           // 1024 entries (CM)
    const i32 *dt2k;         // 256 entries (MATCH)
    const u32 *likely;       // 8 words: bit s set when bit-history state s has seen more ones than zeros
                             // (statetable.v n1 > n0); only steers prefetches, never a result
};

// Model as the kernels see it (pointers are device pointers).
struct ModelDev {
    i32 n, cend, hbegin, hend, header_len;
    u32 h_len, m_len;
    u64 h_off, m_off, r_off, rt_off;  // rt = per-component runtime state (generic kernel)
    u64 ws_bytes;
    const u8 *header;
    const CompDesc *comps;
    // chain-kernel shape
    i32 ctx_mode, n_hash;
    // paged hash tables (chain kernels): comps[].ht_off is then the block's page table (one u32 per
    // kPageBytes of virtual table, 0 = not mapped) and pages come from a pool shared by the wave
    i32 paged;
    u32 pool_pages;
    u8 *pool;
    u32 *pool_next;      // bump allocator
    u32 *pool_overflow;  // set when the pool ran dry: the host repeats the wave with dense tables
};

// Address of the 16-byte slot at virtual offset h0 of a hash table (dense: base + h0; paged: through
// the page table, mapping a zero page on first touch).  Only the owning lane touches its table.
__device__ __forceinline__ u8 *ht_slot(const ModelDev &M, u8 *ht, u32 h0) {
    if (!M.paged) return ht + h0;
    u32 *pt = reinterpret_cast<u32 *>(ht);
    const u32 pg = h0 / kPageBytes;
    u32 pte = pt[pg];
    if (pte == 0) {
        pte = atomicAdd(M.pool_next, 1u) + 1u;
        if (pte > M.pool_pages) {
            *M.pool_overflow = 1u;
            pte = 1u;  // stay in bounds; the result of this wave is discarded
        }
        pt[pg] = pte;
    }
    return M.pool + u64(pte - 1u) * kPageBytes + (h0 & (kPageBytes - 1u));
}

// ---- compression work descriptors ----
struct EncSeg {
    u64 in_off, in_len;    // plaintext range inside d_in
    u64 pay_off, pay_cap;  // coded-payload slot inside the arena
    u32 flags;             // bit0: compress() was called at least once (PP byte is coded, Q16)
    u32 pad;
};
struct EncBlock {
    u32 first_seg, n_seg;
};

// ---- decompression work descriptors ----
struct DecBlock {
    u64 arc_pos;           // archive offset of the first segment marker (just after HCOMP)
    u64 out_off, out_cap;  // plaintext slot inside d_out
};
struct DecBlockOut {
    u64 end_pos;           // archive offset just after the block's 0xFF (or where parsing stopped)
    u64 out_len;           // plaintext bytes produced by the whole block (may exceed out_cap)
    u32 n_seg;
    i32 status;            // ZPAQGPU_* code
};
struct DecSegRec {
    u64 name_off, comment_off;
    u64 out_off, out_len;  // absolute offset inside d_out
    u64 sha_off;           // archive offset of the stored SHA1, ~0 when the marker was not 253
    u32 block, index;
};

// Per-component runtime fields of the generic kernel (Component.a/.b/.c/.cxt/.limit, p[i], h[i]).
struct CompRt {
    i32 a, b, c, limit;
    u32 cxt;
    i32 p;
    u32 h;
    u32 pad;
};

// ---- arithmetic helpers (predictor.v:193-236) ----
__device__ __forceinline__ i32 d_clamp2k(i32 x) { return max(-2048, min(2047, x)); }
__device__ __forceinline__ i32 d_clamp512k(i32 x) { return max(-262144, min(262143, x)); }
// squash index: d+2047 clamped to [0,4093]
__device__ __forceinline__ i32 d_squash_idx(i32 d) { return max(0, min(4093, d + 2047)); }
// stretch index: clamped to [1,32767]
__device__ __forceinline__ i32 d_stretch_idx(i32 p) { return max(1, min(32767, p)); }
// index into stretch_pad for a non-negative argument (entry 0 already holds entry 1)
__device__ __forceinline__ u32 d_stretch_pad_idx(u32 p) { return min(p, 32767u); }

// ---- binary arithmetic coder state (encoder.v:48-89, decoder.v:73-118) ----
// mid = low + ((high-low)*p >> 16) with p < 65536: one IMAD.HI on (p << 16).
__device__ __forceinline__ u32 coder_mid(u32 low, u32 high, u32 p16) {
    return low + __umulhi(high - low, p16 << 16);
}

}  // namespace zg

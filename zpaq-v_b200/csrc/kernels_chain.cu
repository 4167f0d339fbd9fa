// kernels_chain.cu -- the specialised hot-path kernel for models of the shape every predefined
// level has (levels.v:53-375):  ICM -> ISSE -> ... -> ISSE [-> MIX2 of the last two ISSEs], with
// the context hashes coming from one of the two HCOMP programs the levels use.
//
// Mapping (sm_100a):
//  * one ZPAQ block per warp; W warps (blocks) per CTA, one CTA per SM;
//  * squash/stretch/next-state tables (72.5 KiB) live in shared memory, shared by the CTA;
//  * each block's adaptive tables (ICM cm[256], ISSE weight pairs[256] per ISSE: 1 + 2*NI KiB)
//    live in shared memory, private to the warp;
//  * the big hash tables stay in HBM/L2.  A probe (Predictor.find_ht, predictor.v:495-532)
//    touches one 64-byte line; lane i probes for component i with three 16-byte vector loads, the
//    winning 16-byte slot is parked in shared memory for the four bits of the nibble and written
//    back with one 16-byte store at the next nibble boundary;
//  * the bit-serial chain (predict -> code -> update) runs on lane 0 out of registers and shared
//    memory; the coder's low/high/code stay in registers; plaintext/code bytes are staged through
//    shared-memory rings and moved to/from HBM by the whole warp;
//  * MIX2 weights for the 256 possible c8 values of the current byte are staged in shared memory
//    at each byte boundary.
#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"

namespace zg {
namespace {

constexpr int kRing = 256;       // input ring (bytes), power of two
constexpr int kOutStage = 512;   // encoder output stage (bytes)
constexpr u32 kNoSlot = 0xFFFFFFFFu;
constexpr unsigned kFull = 0xFFFFFFFFu;

__host__ __device__ constexpr size_t warp_smem_bytes(int ni, bool mix2) {
    return size_t(256) * 4                 // ICM cm
           + size_t(ni) * 256 * 8          // ISSE weight pairs
           + size_t(ni + 1) * 16           // parked hash slots
           + (mix2 ? 512 : 0)              // staged MIX2 weights
           + kRing + kOutStage;
}
constexpr size_t kSharedTables = 32768 * 2 + 4096 * 2 + 512;

template <int NI, bool MIX2>
struct Chain {
    // shared-memory views
    const int16_t *stretch;
    const u16 *squash;
    const u8 *nex;
    u32 *cm0;
    int2 *wt;
    u8 *slots;   // (NI+1) x 16 bytes
    u16 *a16s;
    u8 *ring;
    u8 *stage;
    // per-lane: the hash table of component `lane`
    u8 *ht;
    u32 ht_len;
    int sizebits;
    u32 slot_at;
    u32 h;        // context hash of component `lane` for the current byte
    // MIX2
    u16 *a16;
    u32 a16_mask, mix_h, mix_sel;
    i32 mix_rate;
    // context history (uniform)
    int ctx_mode, n_hash, n_comp;
    u32 hist;     // CTX_M1: previous three bytes (b1 | b2<<8 | b3<<16); CTX_HASHCHAIN: previous byte
    int lane;

    __device__ void setup(u8 *smem_warp, const ModelDev &M, u8 *ws, const int16_t *st, const u16 *sq,
                          const u8 *nx) {
        lane = threadIdx.x & 31;
        stretch = st, squash = sq, nex = nx;
        u8 *p = smem_warp;
        cm0 = reinterpret_cast<u32 *>(p), p += 1024;
        wt = reinterpret_cast<int2 *>(p), p += size_t(NI) * 2048;
        slots = p, p += (NI + 1) * 16;
        a16s = reinterpret_cast<u16 *>(p), p += MIX2 ? 512 : 0;
        ring = p, p += kRing;
        stage = p;
        ctx_mode = M.ctx_mode, n_hash = M.n_hash, n_comp = M.n;
        // adaptive tables: the fill kernel wrote their initial images into the workspace
        const u32 *src0 = reinterpret_cast<const u32 *>(ws + M.comps[0].cm_off);
        for (int k = lane; k < 256; k += 32) cm0[k] = src0[k];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int2 *src = reinterpret_cast<const int2 *>(ws + M.comps[i + 1].cm_off);
            for (int k = lane; k < 256; k += 32) wt[i * 256 + k] = src[k];
        }
        ht = nullptr, ht_len = 16, sizebits = 0;
        if (lane <= NI) {
            const CompDesc &cd = M.comps[lane];
            ht = ws + cd.ht_off, ht_len = cd.ht_len, sizebits = cd.a + 2;
        }
        slot_at = kNoSlot;
        h = 0, hist = 0, mix_h = 0;
        a16 = nullptr, a16_mask = 0, mix_sel = 0, mix_rate = 0;
        if (MIX2) {
            const CompDesc &cd = M.comps[NI + 1];
            a16 = reinterpret_cast<u16 *>(ws + cd.a16_off);
            a16_mask = cd.a16_len - 1, mix_sel = cd.p[3], mix_rate = i32(cd.p[2]);
        }
        __syncwarp();
    }

    // pr.reset() (predictor.v:827-833): contexts go to zero, history and tables stay.
    __device__ void segment_reset() {
        h = 0, mix_h = 0;
        stage_mix();
    }

    __device__ void stage_mix() {
        if (MIX2) {
            __syncwarp();
            for (int k = lane; k < 256; k += 32) a16s[k] = a16[(mix_h + u32(k)) & a16_mask];
            __syncwarp();
        }
    }

    // After byte c: run the HCOMP program in closed form (levels.v:72-87, :126-139) and latch the
    // hash of this lane's component (predictor.v:809-818).
    __device__ void byte_end(u32 c) {
        u32 mine = 0, mixv = 0;
        if (ctx_mode == CTX_M1) {
            u32 a = (0u + c + 512u) * 773u;
            a = (a + (hist & 255u) + 512u) * 773u;
            const u32 h0 = a;
            a = (a + ((hist >> 8) & 255u) + 512u) * 773u;
            a = (a + ((hist >> 16) & 255u) + 512u) * 773u;
            mine = lane == 0 ? h0 : (lane == 1 ? a : 0u);
            hist = ((hist << 8) | c) & 0xFFFFFFu;
        } else {
            u32 a = c;
            const u32 q = hist;
            for (int r = 0; r < n_hash; ++r) {
                a = (a + q + 512u) * 773u;
                if (r == lane) mine = a;
                if (MIX2 && r == NI + 1) mixv = a;
            }
            hist = c;
        }
        h = lane < n_comp ? mine : 0u;
        if (MIX2) {
            mix_h = mixv;
            stage_mix();
        }
    }

    // Predictor.find_ht for component `lane` (predictor.v:495-532), whole warp converged.
    __device__ void probe(u32 c8v) {
        __syncwarp();
        if (lane <= NI) {
            uint4 *park = reinterpret_cast<uint4 *>(slots) + lane;
            if (slot_at != kNoSlot) *reinterpret_cast<uint4 *>(ht + slot_at) = *park;
            const u32 key = h + 16u * c8v;
            const u32 chk = (key >> sizebits) & 255u;
            const u32 h0 = (key * 16u) & (ht_len - 16u), h1 = h0 ^ 16u, h2 = h0 ^ 32u;
            const uint4 s0 = *reinterpret_cast<const uint4 *>(ht + h0);
            const uint4 s1 = *reinterpret_cast<const uint4 *>(ht + h1);
            const uint4 s2 = *reinterpret_cast<const uint4 *>(ht + h2);
            uint4 pick;
            u32 at;
            if ((s0.x & 255u) == chk) {
                pick = s0, at = h0;
            } else if ((s1.x & 255u) == chk) {
                pick = s1, at = h1;
            } else if ((s2.x & 255u) == chk) {
                pick = s2, at = h2;
            } else {
                const u32 q0 = (s0.x >> 8) & 255u, q1 = (s1.x >> 8) & 255u, q2 = (s2.x >> 8) & 255u;
                at = (q0 <= q1 && q0 <= q2) ? h0 : (q1 < q2 ? h1 : h2);
                pick = make_uint4(chk, 0u, 0u, 0u);
            }
            *park = pick;
            slot_at = at;
        }
        __syncwarp();
    }
};

// Four coded bits of one nibble on lane 0.  ENC: `nib` holds the 4 bits, MSB first.
// Returns the nibble (decoder) / nib (encoder).  c8 enters as 1 (high nibble) or 16..31.
template <int NI, bool MIX2, bool DEC, class IO>
__device__ __forceinline__ u32 code_nibble(Chain<NI, MIX2> &C, u32 nib, u32 &c8, u32 &low, u32 &high,
                                           u32 &code, IO &io) {
    u32 idx = 1;
    u32 got = 0;
#pragma unroll 1
    for (int k = 3; k >= 0; --k) {
        // ---- predict (predictor.v:555-563, :615-631, :586-599) ----
        i32 p[NI + 2];
        u32 st[NI + 1];
        int2 w[NI + 1];
        st[0] = C.slots[idx];
        const u32 v0 = C.cm0[st[0]];
        p[0] = C.stretch[d_stretch_idx(i32(v0 >> 8))];
#pragma unroll
        for (int i = 1; i <= NI; ++i) {
            st[i] = C.slots[i * 16 + idx];
            w[i] = C.wt[(i - 1) * 256 + st[i]];
        }
#pragma unroll
        for (int i = 1; i <= NI; ++i) p[i] = d_clamp2k((w[i].x * p[i - 1] + w[i].y * 64) >> 16);
        i32 mw = 0;
        u32 msel = 0;
        if (MIX2) {
            msel = c8 & C.mix_sel;
            mw = C.a16s[msel];
            p[NI + 1] = d_clamp2k((mw * p[NI - 1] + (65536 - mw) * p[NI]) >> 16);
        }
        const i32 pl = MIX2 ? p[NI + 1] : p[NI];
        const u32 p16 = u32(C.squash[d_squash_idx(pl)]) * 2u + 1u;
        // ---- code (encoder.v:48-89 / decoder.v:73-118) ----
        const u32 mid = coder_mid(low, high, p16);
        u32 y;
        if (DEC) {
            y = code <= mid;
        } else {
            y = (nib >> k) & 1u;
        }
        if (y) high = mid; else low = mid + 1;
        while ((high ^ low) < 0x1000000u) {
            if (!DEC) io.put(high >> 24);
            low <<= 8;
            high = (high << 8) | 0xFFu;
            if (low == 0) low = 1;
            if (DEC) code = (code << 8) | io.get();
        }
        // ---- update (predictor.v:701-709, :776-791, :744-762) ----
        const i32 t = y ? 32767 : 0;
        C.slots[idx] = C.nex[st[0] * 2 + y];
        C.cm0[st[0]] = u32(i32(v0) + ((t - i32(v0 >> 8)) >> 2));
#pragma unroll
        for (int i = 1; i <= NI; ++i) {
            const i32 err = t - i32(C.squash[d_squash_idx(p[i])]);
            int2 nw;
            nw.x = d_clamp512k(w[i].x + ((err * p[i - 1] + 4096) >> 13));
            nw.y = d_clamp512k(w[i].y + ((err + 16) >> 5));
            C.wt[(i - 1) * 256 + st[i]] = nw;
            C.slots[i * 16 + idx] = C.nex[st[i] * 2 + y];
        }
        if (MIX2) {
            const i32 err = ((t - i32(p16 >> 1)) * C.mix_rate) >> 5;
            i32 nw = mw + ((err * (p[NI - 1] - p[NI]) + 4096) >> 13);
            nw = max(0, min(65535, nw));
            C.a16s[msel] = u16(nw);
            C.a16[(C.mix_h + msel) & C.a16_mask] = u16(nw);
        }
        c8 = (c8 << 1) | y;
        idx = (idx * 2 + y) & 15u;
        got = (got << 1) | y;
    }
    return got;
}

// ---- lane-0 byte I/O over the shared-memory stages ----
struct EncIO {
    u8 *stage;
    u32 fill;
    __device__ __forceinline__ void put(u32 b) { stage[fill++] = u8(b); }
    __device__ __forceinline__ u32 get() { return 0; }
};
struct DecIO {
    const u8 *ring;
    u64 pos;
    __device__ __forceinline__ void put(u32) {}
    __device__ __forceinline__ u32 get() { return ring[(pos++) & (kRing - 1)]; }
};

// Keep at least `need` bytes of [pos, limit) in the ring; bytes past `limit` read as zero (the
// reference's get() returns -1 there and the decoder shifts in nothing, decoder.v:111-116).
__device__ __forceinline__ void ring_fill(u8 *ring, const u8 *base, u64 pos, u64 &filled, u64 limit,
                                          int lane) {
    if (filled < pos + 64) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u64 at = filled + u64(lane + 32 * k);
            ring[at & (kRing - 1)] = at < limit ? base[at] : u8(0);
        }
        filled += 128;
        __syncwarp();
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// k_encode_chain
// ------------------------------------------------------------------------------------------
template <int NI, bool MIX2>
__global__ void __launch_bounds__(256, 1) k_encode_chain(EncodeArgs A) {
    extern __shared__ __align__(16) u8 smem[];
    int16_t *s_stretch = reinterpret_cast<int16_t *>(smem);
    u16 *s_squash = reinterpret_cast<u16 *>(smem + 65536);
    u8 *s_nex = smem + 65536 + 8192;
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(A.tables.stretch);
        uint4 *d = reinterpret_cast<uint4 *>(s_stretch);
        for (int k = threadIdx.x; k < 4096; k += blockDim.x) d[k] = g[k];
        const uint4 *g2 = reinterpret_cast<const uint4 *>(A.tables.squash);
        uint4 *d2 = reinterpret_cast<uint4 *>(s_squash);
        for (int k = threadIdx.x; k < 512; k += blockDim.x) d2[k] = g2[k];
        for (int k = threadIdx.x; k < 512; k += blockDim.x) s_nex[k] = A.tables.nex[k];
    }
    __syncthreads();
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + wic;
    if (slot >= A.n_blocks) return;
    Chain<NI, MIX2> C;
    u8 *ws = A.workspace + u64(slot) * A.model.ws_bytes;
    C.setup(smem + kSharedTables + size_t(wic) * warp_smem_bytes(NI, MIX2), A.model, ws, s_stretch,
            s_squash, s_nex);
    const EncBlock blk = A.blocks[A.first_block + slot];
    for (u32 s = 0; s < blk.n_seg; ++s) {
        const EncSeg seg = A.segs[blk.first_seg + s];
        const u8 *src = A.in + seg.in_off;
        u8 *dst = A.arena + seg.pay_off;
        C.segment_reset();
        u32 low = 1, high = 0xFFFFFFFFu, code = 0;
        EncIO io{C.stage, 0};
        u64 written = 0;   // payload bytes already moved (or counted) past the stage
        u64 filled = 0;    // ring holds plaintext [.., filled)
        const bool pp = (seg.flags & 1u) != 0;
        const u64 total = seg.in_len + (pp ? 1 : 0);
        for (u64 k = 0; k < total; ++k) {
            u32 c = 0;
            if (!(pp && k == 0)) {
                const u64 at = pp ? k - 1 : k;
                ring_fill(C.ring, src, at, filled, seg.in_len, lane);
                c = C.ring[at & (kRing - 1)];
            }
            C.probe(1u);
            u32 c8 = 1;
            if (lane == 0) {
                // "not EOF" flag: encode(0, p=0) => low += 1 (encoder.v:108)
                low = low + 1;
                while ((high ^ low) < 0x1000000u) {
                    io.put(high >> 24);
                    low <<= 8;
                    high = (high << 8) | 0xFFu;
                    if (low == 0) low = 1;
                }
                code_nibble<NI, MIX2, false>(C, c >> 4, c8, low, high, code, io);
            }
            C.probe(16u | (c >> 4));
            if (lane == 0) {
                c8 = 16u | (c >> 4);
                code_nibble<NI, MIX2, false>(C, c & 15u, c8, low, high, code, io);
            }
            C.byte_end(c);
            // move full 256-byte chunks of coded output to HBM
            const u32 fill = __shfl_sync(kFull, io.fill, 0);
            if (fill >= 256) {
                __syncwarp();
                u32 tail = 0;
                if (lane + 256 < int(fill)) tail = C.stage[256 + lane];
                for (int q = lane; q < 256; q += 32)
                    if (written + q < seg.pay_cap) dst[written + q] = C.stage[q];
                __syncwarp();
                if (lane + 256 < int(fill)) C.stage[lane] = u8(tail);
                // a byte can emit at most 36 coded bytes, so the tail beyond 256 fits 32 lanes + 4
                for (int q = 288 + lane; q < int(fill); q += 32) C.stage[q - 256] = C.stage[q];
                __syncwarp();
                written += 256;
                io.fill = fill - 256;
            }
        }
        // EOF: encode(1, p=0) then flush the four bytes of high (encoder.v:101-105, :130-139)
        if (lane == 0) {
            high = low;
            while ((high ^ low) < 0x1000000u) {
                io.put(high >> 24);
                low <<= 8;
                high = (high << 8) | 0xFFu;
                if (low == 0) low = 1;
            }
            io.put(high >> 24), io.put((high >> 16) & 255u), io.put((high >> 8) & 255u), io.put(high & 255u);
        }
        const u32 fill = __shfl_sync(kFull, io.fill, 0);
        __syncwarp();
        for (u32 q = lane; q < fill; q += 32)
            if (written + q < seg.pay_cap) dst[written + q] = C.stage[q];
        __syncwarp();
        if (lane == 0) A.pay_len[blk.first_seg + s] = written + fill;
    }
}

// ------------------------------------------------------------------------------------------
// k_decode_chain
// ------------------------------------------------------------------------------------------
template <int NI, bool MIX2>
__global__ void __launch_bounds__(256, 1) k_decode_chain(DecodeArgs A) {
    extern __shared__ __align__(16) u8 smem[];
    int16_t *s_stretch = reinterpret_cast<int16_t *>(smem);
    u16 *s_squash = reinterpret_cast<u16 *>(smem + 65536);
    u8 *s_nex = smem + 65536 + 8192;
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(A.tables.stretch);
        uint4 *d = reinterpret_cast<uint4 *>(s_stretch);
        for (int k = threadIdx.x; k < 4096; k += blockDim.x) d[k] = g[k];
        const uint4 *g2 = reinterpret_cast<const uint4 *>(A.tables.squash);
        uint4 *d2 = reinterpret_cast<uint4 *>(s_squash);
        for (int k = threadIdx.x; k < 512; k += blockDim.x) d2[k] = g2[k];
        for (int k = threadIdx.x; k < 512; k += blockDim.x) s_nex[k] = A.tables.nex[k];
    }
    __syncthreads();
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + wic;
    if (slot >= A.n_blocks) return;
    const int bi = A.first_block + slot;
    Chain<NI, MIX2> C;
    u8 *ws = A.workspace + u64(slot) * A.model.ws_bytes;
    C.setup(smem + kSharedTables + size_t(wic) * warp_smem_bytes(NI, MIX2), A.model, ws, s_stretch,
            s_squash, s_nex);
    const DecBlock blk = A.blocks[bi];
    const u8 *arc = A.arc;
    u64 pos = blk.arc_pos;  // uniform across the warp
    DecBlockOut res;
    res.end_pos = pos, res.out_len = 0, res.n_seg = 0, res.status = ZPAQGPU_OK;
    u8 *dst = A.out + blk.out_off;
    auto rd = [&](u64 at) -> i32 { return at < A.arc_len ? i32(arc[at]) : -1; };
    for (;;) {
        const i32 marker = rd(pos++);  // decompressor.v:356-365
        if (marker < 0) { res.status = ZPAQGPU_E_FORMAT; break; }
        if (marker == 0xFF) break;
        DecSegRec rec;
        rec.block = u32(bi), rec.index = res.n_seg, rec.sha_off = ~0ull;
        rec.name_off = pos;
        i32 c;
        bool block_over = false;
        while ((c = rd(pos++)) > 0)
            if (c == 0xFF) { block_over = true; break; }
        if (block_over) break;
        if (c < 0) { res.status = ZPAQGPU_E_FORMAT; break; }
        rec.comment_off = pos;
        while ((c = rd(pos++)) > 0) {}
        if (c < 0 || rd(pos++) < 0) { res.status = ZPAQGPU_E_FORMAT; break; }
        C.segment_reset();
        // Decoder.init (decoder.v:29-47)
        u32 low = 1, high = 0xFFFFFFFFu, code = 0;
        u64 filled = pos;
        ring_fill(C.ring, arc, pos, filled, A.arc_len, lane);
        DecIO io{C.ring, pos};
        for (int k = 0; k < 4; ++k) code = (code << 8) | io.get();
        rec.out_off = blk.out_off + res.out_len;
        u64 produced = 0;   // plaintext bytes of this segment
        u32 staged = 0;     // of which still in the stage
        int pp_state = 0;
        bool unsupported = false;
        for (;;) {
            ring_fill(C.ring, arc, io.pos, filled, A.arc_len, lane);
            u32 c8 = 1, eof = 0;
            if (lane == 0) {
                // EOF flag: decode(p=0) => y = (code <= low) (decoder.v:128-131)
                eof = code <= low;
                if (eof) high = low; else low = low + 1;
                while ((high ^ low) < 0x1000000u) {
                    low <<= 8;
                    high = (high << 8) | 0xFFu;
                    if (low == 0) low = 1;
                    code = (code << 8) | io.get();
                }
            }
            eof = __shfl_sync(kFull, eof, 0);
            if (eof) break;
            // the model is only consulted once a data byte is known to follow (decoder.v:128-142):
            // a probe at EOF could evict a slot that a later segment of the block still needs
            C.probe(1u);
            if (lane == 0) code_nibble<NI, MIX2, true>(C, 0, c8, low, high, code, io);
            c8 = __shfl_sync(kFull, c8, 0);
            C.probe(c8);
            if (lane == 0) code_nibble<NI, MIX2, true>(C, 0, c8, low, high, code, io);
            const u32 ch = __shfl_sync(kFull, c8, 0) & 255u;
            io.pos = __shfl_sync(kFull, io.pos, 0);
            C.byte_end(ch);
            if (pp_state == 0) {  // PostProcessor.write state 0 (decompressor.v:58-70)
                pp_state = ch == 1 ? 2 : 1;
                if (pp_state == 2) { unsupported = true; break; }
            } else {
                if (lane == 0) C.stage[staged] = u8(ch);
                ++staged, ++produced;
                if (staged == 256) {
                    __syncwarp();
                    const u64 base = res.out_len + produced - 256;
                    for (int q = lane; q < 256; q += 32)
                        if (base + q < blk.out_cap) dst[base + q] = C.stage[q];
                    __syncwarp();
                    staged = 0;
                }
            }
        }
        io.pos = __shfl_sync(kFull, io.pos, 0);
        code = __shfl_sync(kFull, code, 0);
        if (unsupported) { res.status = ZPAQGPU_E_UNSUPPORTED; break; }
        if (staged) {
            __syncwarp();
            const u64 base = res.out_len + produced - staged;
            for (u32 q = lane; q < staged; q += 32)
                if (base + q < blk.out_cap) dst[base + q] = C.stage[q];
            __syncwarp();
        }
        // Decoder.skip (decoder.v:151-196) and read_segment_end (decompressor.v:608-631)
        pos = io.pos;
        u32 curr = code;
        i32 mk = 0;
        bool eofs = false;
        if (curr == 0) {
            const i32 b = rd(pos++);
            if (b < 0) eofs = true; else curr = u32(b);
        }
        while (!eofs && curr != 0) {
            const i32 b = rd(pos++);
            if (b < 0) eofs = true; else curr = (curr << 8) | u32(b);
        }
        while (!eofs) {
            mk = rd(pos++);
            if (mk < 0) eofs = true;
            if (mk != 0) break;
        }
        if (!eofs && mk == 253) {
            rec.sha_off = pos;
            pos = min(pos + 20, A.arc_len);
        }
        if (pos > A.arc_len) pos = A.arc_len;
        rec.out_len = produced;
        res.out_len += produced;
        if (lane == 0) {
            const u32 at = atomicAdd(A.seg_count, 1u);
            if (at < A.seg_cap) A.seg_recs[at] = rec;
        }
        res.n_seg++;
    }
    if (pos > A.arc_len) pos = A.arc_len;
    res.end_pos = pos;
    if (lane == 0) A.results[bi] = res;
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
size_t chain_smem_bytes(const Model &m, int warps_per_cta) {
    return kSharedTables + size_t(warps_per_cta) * warp_smem_bytes(m.n_isse, m.has_mix2);
}
int chain_max_warps_per_cta(const Model &m) {
    const size_t budget = 227 * 1024;
    const size_t per = warp_smem_bytes(m.n_isse, m.has_mix2);
    int w = int((budget - kSharedTables) / per);
    return w > 8 ? 8 : w;
}

template <int NI, bool MIX2>
static bool launch_pair(bool decode, const EncodeArgs *E, const DecodeArgs *D, int n_blocks, int wpc,
                        size_t smem, cudaStream_t s) {
    const int grid = (n_blocks + wpc - 1) / wpc;
    if (decode) {
        auto k = k_decode_chain<NI, MIX2>;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
            return false;
        k<<<grid, wpc * 32, smem, s>>>(*D);
    } else {
        auto k = k_encode_chain<NI, MIX2>;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
            return false;
        k<<<grid, wpc * 32, smem, s>>>(*E);
    }
    return true;
}

static bool dispatch(const Model &m, bool decode, const EncodeArgs *E, const DecodeArgs *D, int n_blocks,
                     int wpc, cudaStream_t s) {
    if (!m.is_chain) return false;
    const size_t smem = chain_smem_bytes(m, wpc);
#define ZG_CASE(NI, MX) \
    if (m.n_isse == NI && m.has_mix2 == MX) return launch_pair<NI, MX>(decode, E, D, n_blocks, wpc, smem, s);
    ZG_CASE(0, false) ZG_CASE(1, false) ZG_CASE(2, false) ZG_CASE(3, false) ZG_CASE(4, false)
    ZG_CASE(5, false) ZG_CASE(6, false) ZG_CASE(7, false)
    ZG_CASE(2, true) ZG_CASE(3, true) ZG_CASE(4, true) ZG_CASE(5, true) ZG_CASE(6, true) ZG_CASE(7, true)
#undef ZG_CASE
    return false;
}

bool launch_encode_chain(const Model &m, const EncodeArgs &A, int wpc, cudaStream_t s) {
    return dispatch(m, false, &A, nullptr, A.n_blocks, wpc, s);
}
bool launch_decode_chain(const Model &m, const DecodeArgs &A, int wpc, cudaStream_t s) {
    return dispatch(m, true, nullptr, &A, A.n_blocks, wpc, s);
}

}  // namespace zg

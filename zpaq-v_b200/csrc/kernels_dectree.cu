// kernels_dectree.cu -- k_decode_tree2: the decompressor's hot kernel for -m1..-m3 shaped models
// (ICM -> NI x ISSE, no MIX2), TWO WARPS PER ZPAQ BLOCK.
//
// A block is one bit-serial chain; with 1 024 blocks on 592 warp schedulers the GPU has issue slots to
// spare and every chain is bound by the in-order instruction stream of its own warp (round 1: 540
// instructions per nibble at 4.3 cycles each).  The work of a block is therefore split by data
// ownership between two warps on different schedulers:
//
//   R (rounds)  owns the adaptive tables (ICM {cm, stretch(cm>>8)}, ISSE {wt0, wt1}, shared memory) and
//               the arithmetic decoder.  Per nibble: lane L = (tree node L & 15, assumed outcome L >> 4)
//               reads its node's bit-history states from the slot images S published, evaluates the
//               chain in four lock-step rounds (one per tree level; an ancestor's update reaches the
//               descendants that share its state by SHFL), walks the tree with the decoder, and the
//               nodes on the decoded path store their updates.  R never touches a hash table.
//   S (slots)   owns the hash tables in HBM (Predictor.find_ht, predictor.v:495-532) and byte I/O.  After
//               two decoded bits the slot of the next nibble is one of four: lane (component L & 7,
//               candidate L >> 3) requests the three 16-byte candidate slots of its line, makes the
//               find_ht choice for ITS candidate from registers and publishes the chosen slot image, so
//               that R finds the next nibble's states waiting when it gets there.  When the nibble is
//               complete S inserts the successor states on the decoded path into the current slot,
//               writes it back (one 16-byte store), latches the context hashes of the next byte
//               (closed forms of the level programs, levels.v:72-87, :126-139) and stages the plaintext.
//
// Hand-offs: S -> R through one named barrier per block (S: bar.arrive after st.shared, R: bar.sync --
// the producer/consumer pattern of the PTX manual), R -> S through mbarriers in shared memory (R: one
// elected st.shared + mbarrier.arrive, never waits; S: mbarrier.try_wait sleeps in hardware).  Every
// wait has exactly one matching arrival in program order of the other warp (see "protocol" below), so
// the schedule cannot deadlock or lap.  Results are bit-identical to the serial evaluation
// (predictor.v:536-824): same table reads, same updates, in the same order.
#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"

#include <cstdio>

namespace zg {
namespace {

#ifdef ZG_TIMING
__device__ long long g_ts[8];      // timestamps exchanged between the two warps of slot 0
__device__ long long g_acc[16];    // accumulated phase times of slot 0
#define ZT(stmt) do { if (tslot0) { stmt; } } while (0)
#else
#define ZT(stmt) do { } while (0)
#endif

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kRing = 256;      // code-byte ring of R
constexpr int kOutStage = 256;  // plaintext stage of S
constexpr size_t kSharedTables = 32768 * 2 + 4096 * 2 + 512 + 64;  // stretch, squash, next-state pairs, likely bits

// control codes on the `part` mailbox (data values are 0..3)
constexpr u32 kSegBegin = 0x100u, kSegEnd = 0x101u, kExit = 0x102u;

__host__ __device__ constexpr size_t pair_smem_bytes(int ni) {
    return size_t(ni + 1) * 2048          // adaptive tables (R)
           + 4 * 8 * 16                   // slot images: candidate x component (S -> R)
           + 64                           // mailboxes and mbarriers
           + kRing + kOutStage;
}

__device__ __forceinline__ uint4 ldg128(const u8 *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 smem_addr(const void *p) { return u32(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(u32 addr, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try(u32 addr, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait sleeps in hardware until the phase completes or a system time limit passes.  A wait that
// lasts a second means the two warps of a block lost step: trap (the launch fails loudly) instead of
// hanging the device.
__device__ __forceinline__ void mbar_wait(u32 addr, u32 parity) {
    if (mbar_try(addr, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(addr, parity))
        if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// Shared-memory views of one block (both warps).
struct PairMem {
    int2 *tabs;        // NC x 256
    u8 *cand;          // [candidate 0..3][component 0..7][16]
    volatile u32 *msg_part;
    volatile u32 *msg_full;  // [2]
    u32 mb_part, mb_full0, mb_full1;  // shared-space addresses
    u32 pull_sink;                    // shared-space address of 16 bytes nobody reads
    u8 *ring, *stage;
};

template <int NI>
__device__ __forceinline__ PairMem carve_pair(u8 *p) {
    PairMem m;
    m.tabs = reinterpret_cast<int2 *>(p), p += size_t(NI + 1) * 2048;
    m.cand = p, p += 4 * 8 * 16;
    u64 *bars = reinterpret_cast<u64 *>(p);
    m.mb_part = smem_addr(bars), m.mb_full0 = smem_addr(bars + 1), m.mb_full1 = smem_addr(bars + 2);
    m.msg_part = reinterpret_cast<volatile u32 *>(p + 32);
    m.msg_full = reinterpret_cast<volatile u32 *>(p + 40);
    m.pull_sink = smem_addr(p + 48);
    p += 64;
    m.ring = p, p += kRing;
    m.stage = p;
    return m;
}

// HCOMP in closed form (levels.v:72-87, :126-139): the context hash component `sel` gets for the
// byte that follows byte c, given the history before c.
struct CtxHash {
    int mode, n_hash, n_comp;
    u32 hist;  // CTX_M1: previous three bytes; CTX_HASHCHAIN: previous byte
    __device__ __forceinline__ u32 next(u32 c, int sel) const {
        u32 mine = 0;
        if (mode == CTX_M1) {
            u32 a = (0u + c + 512u) * 773u;
            a = (a + (hist & 255u) + 512u) * 773u;
            const u32 h0 = a;
            a = (a + ((hist >> 8) & 255u) + 512u) * 773u;
            a = (a + ((hist >> 16) & 255u) + 512u) * 773u;
            mine = sel == 0 ? h0 : (sel == 1 ? a : 0u);
        } else {
            u32 a = c;
            for (int r = 0; r < n_hash; ++r) {
                a = (a + hist + 512u) * 773u;
                if (r == sel) mine = a;
            }
        }
        return sel < n_comp ? mine : 0u;
    }
    __device__ __forceinline__ void advance(u32 c) { hist = mode == CTX_M1 ? (((hist << 8) | c) & 0xFFFFFFu) : c; }
};

// Predictor.find_ht's choice among the three candidate slots of a line (predictor.v:495-532): a slot
// whose check byte matches, else the one with the lowest priority byte (ties: first, then third unless
// the second is lower) cleared and tagged.  Selects only, no branches.
__device__ __forceinline__ void pick_slot(u32 chk, u8 *b0, const uint4 &s0, const uint4 &s1, const uint4 &s2,
                                          uint4 &chosen, u8 *&at) {
    u8 *b1 = reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 16u);
    u8 *b2 = reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 32u);
    const bool m0 = (s0.x & 255u) == chk, m1 = (s1.x & 255u) == chk, m2 = (s2.x & 255u) == chk;
    const u32 p0 = (s0.x >> 8) & 255u, p1 = (s1.x >> 8) & 255u, p2 = (s2.x >> 8) & 255u;
    u8 *victim = (p0 <= p1 && p0 <= p2) ? b0 : (p1 < p2 ? b1 : b2);
    const bool hit = m0 | m1 | m2;
    at = m0 ? b0 : m1 ? b1 : m2 ? b2 : victim;
    const uint4 pick = m0 ? s0 : (m1 ? s1 : s2);
    chosen.x = hit ? pick.x : chk;
    chosen.y = hit ? pick.y : 0u;
    chosen.z = hit ? pick.z : 0u;
    chosen.w = hit ? pick.w : 0u;
}

// ------------------------------------------------------------------------------------------
// S: hash slots, context hashes, plaintext staging
// ------------------------------------------------------------------------------------------
template <int NI>
__device__ void run_slots(const DecodeArgs &A, const PairMem &P, int bar_id, int slot, int bi, const u16 *nex16,
                          const u32 *likely) {
    const ModelDev &M = A.model;
    const int lane = threadIdx.x & 31;
    const int pc = lane & 7;
    const u32 pcand = u32(lane) >> 3;
    const bool powner = pc <= NI;
    u8 *ws = A.workspace + u64(slot) * M.ws_bytes;
    u8 *ht = nullptr;
    u32 ht_len = 16;
    int sizebits = 0;
    if (powner) {
        const CompDesc &cd = M.comps[pc];
        ht = ws + cd.ht_off, ht_len = cd.ht_len, sizebits = cd.a + 2;
    }
    const bool spec = (A.flags & 1) != 0;
    const u32 n_guess = spec ? (u32(A.flags) >> 8) & 7u : 0u;
#ifdef ZG_GUESS_STATS
    u32 g_last = 0xFFu, g_hit = 0, g_tot = 0, g_none = 0;
#endif
    const DecBlock blk = A.blocks[bi];
    u8 *dst = A.out + blk.out_off;
    u64 out_done = 0;  // plaintext bytes of earlier segments of the block

    CtxHash cx{M.ctx_mode, M.n_hash, M.n, 0u};
    u32 h = 0;          // context hash of component pc for the current byte (predictor.v:813-815)
    u32 c8base = 1;     // c8 at the start of the current nibble: 1 or 16 | high nibble
    // the slot of the current nibble (acting lane of component pc: image and address) ...
    uint4 sl = make_uint4(0, 0, 0, 0);
    u8 *slot_at = nullptr;
    bool actor = false;
    u32 cur_vline = ~0u;
    // ... and the slot of the nibble before, whose write-back waits for the shadow of the next loads
    uint4 psl = make_uint4(0, 0, 0, 0);
    u8 *pslot_at = nullptr;
    bool pending = false;   // this lane holds a slot to write back
    u32 pf = 0;             // the decoded nibble whose successor states go into it
    u32 prev_vline = ~0u;
    u32 ph_part = 0, ph_full0 = 0, ph_full1 = 0, k = 0;
    // per segment
    u64 produced = 0;
    u32 staged = 0;
    int pp_state = 0;
    u32 byte_done = 0x100u;  // byte completed by the last nibble, not yet staged
#ifdef ZG_TIMING
    const bool tslot0 = slot == 0 && lane == 0;
    long long ts0 = 0, ts1 = 0, ts2 = 0;
#endif

    auto slot_peek = [&](u32 h0) -> u8 * {
        if (!M.paged) return ht + h0;
        const u32 pte = reinterpret_cast<const u32 *>(ht)[h0 / kPageBytes];
        return pte ? M.pool + u64(pte - 1u) * kPageBytes + (h0 & (kPageBytes - 1u)) : nullptr;
    };
    // key of the slot a nibble probes when c8 (with its leading one) has grown to c8new: the low nibble of
    // the same byte (c8new 16..31) or the high nibble of the next byte (c8new >= 256)
    auto key_of = [&](u32 c8new) -> u32 { return c8new < 256u ? h + 16u * c8new : cx.next(c8new & 255u, pc) + 16u; };
    // find_ht now, from memory, for the nibble that starts with c8base
    auto fresh = [&](u32 cand, uint4 &img, u8 *&at) {
        const u32 key = h + 16u * c8base;
        const u32 h0 = (key * 16u) & (ht_len - 16u);
        u8 *b0 = ht_slot(M, ht, h0);
        const uint4 s0 = ldg128(b0);
        const uint4 s1 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 16u));
        const uint4 s2 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 32u));
        pick_slot((key >> sizebits) & 255u, b0, s0, s1, s2, img, at);
        *reinterpret_cast<uint4 *>(P.cand + (cand * 8u + u32(pc)) * 16u) = img;
    };
    // successor states on the decoded path into the slot (statetable.v:75-88), slot back to its table
    auto write_back = [&](bool mine, uint4 img, u8 *at, u32 f) {
        if (mine) {
            u32 idx = 1;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const u32 y = (f >> (3 - b)) & 1u;
                const u32 sh = (b == 0) ? 8u : (idx & 3u) * 8u;
                const bool hiword = (b == 3) && (idx & 4u);
                const u32 word = (b < 2) ? img.x : (b == 2) ? img.y : (hiword ? img.w : img.z);
                const u32 st = (word >> sh) & 255u;
                const u32 ns = (u32(nex16[st]) >> (y * 8u)) & 255u;
                const u32 d = (st ^ ns) << sh;
                if (b < 2) img.x ^= d;
                else if (b == 2) img.y ^= d;
                else if (hiword) img.w ^= d;
                else img.z ^= d;
                idx = (idx * 2 + y) & 15u;
            }
            *reinterpret_cast<uint4 *>(at) = img;
        }
    };
    auto flush_stage = [&](u32 n) {
        __syncwarp();
        const u64 base = out_done + produced - n;
        for (u32 q = lane; q < n; q += 32)
            if (base + q < blk.out_cap) dst[base + q] = P.stage[q];
        __syncwarp();
    };
    auto stage_byte = [&]() {
        if (byte_done < 0x100u) {
            if (pp_state == 0) {  // PostProcessor.write state 0 (decompressor.v:58-70)
                pp_state = byte_done == 1u ? 2 : 1;
            } else if (pp_state == 1) {
                if (lane == 0) P.stage[staged] = u8(byte_done);
                ++staged, ++produced;
                if (staged == u32(kOutStage)) {
                    flush_stage(staged);
                    staged = 0;
                }
            }
            byte_done = 0x100u;
        }
    };

    for (;;) {
        mbar_wait(P.mb_part, ph_part), ph_part ^= 1u;
        const u32 m = *P.msg_part;
        if (m >= kSegBegin) {
            // whatever is still owed to the tables and the output goes out first
            write_back(pending, psl, pslot_at, pf);
            pending = false, prev_vline = ~0u;
            __syncwarp();
            stage_byte();
            if (m == kExit) break;
            if (m == kSegBegin) {
                // pr.reset() (predictor.v:827-833): contexts to zero, history and tables stay
                h = 0, c8base = 1;
                produced = 0, staged = 0, pp_state = 0;
                actor = powner && pcand == 0u;
                if (actor) fresh(0u, sl, slot_at);
                cur_vline = ((((h + 16u * c8base) * 16u) & (ht_len - 16u)) >> 6);
            } else {  // kSegEnd
                if (staged) flush_stage(staged);
                out_done += produced;
            }
            pair_arrive(bar_id);
            continue;
        }
        ZT(ts0 = clock64(); g_acc[6] += ts0 - g_ts[0]);
        // ---- two bits of the current nibble are known: request the four possible next slots ----
        bool q_ok = false;
        u32 q_chk = 0;
        u8 *qb0 = nullptr;
        uint4 s0 = make_uint4(0, 0, 0, 0), s1 = s0, s2 = s0;
        if (powner && spec) {
            const u32 q_key = key_of((((c8base << 2) | m) << 2) | pcand);
            const u32 h0 = (q_key * 16u) & (ht_len - 16u);
            qb0 = slot_peek(h0);
            q_chk = (q_key >> sizebits) & 255u;
            // The line of the slot in use changes at its write-back, the line of the slot before is about
            // to (below): neither is requested early.
            if (qb0 && (h0 >> 6) != cur_vline && (h0 >> 6) != prev_vline) {
                s0 = ldg128(qb0);
                s1 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(qb0) ^ 16u));
                s2 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(qb0) ^ 32u));
                q_ok = true;
            }
        }
        // ---- in the shadow of those loads: the slot of the previous nibble goes back to its table, the
        //      byte it completed to the plaintext stage ----
        write_back(pending, psl, pslot_at, pf);
        pending = false, prev_vline = ~0u;
        stage_byte();
        // ---- find_ht's choice for every candidate, published for R ----
        uint4 qsl = make_uint4(0, 0, 0, 0);
        u8 *qat = nullptr;
        if (q_ok) {
            pick_slot(q_chk, qb0, s0, s1, s2, qsl, qat);
            *reinterpret_cast<uint4 *>(P.cand + (pcand * 8u + u32(pc)) * 16u) = qsl;
        }
        const bool all_ok = __all_sync(kFull, q_ok || !powner);
        ZT(ts1 = clock64(); g_acc[7] += ts1 - ts0; if (all_ok) g_ts[2] = ts1; else g_acc[12] += 1);
        if (all_ok) pair_arrive(bar_id);  // whichever candidate wins, its images are in place
        // ---- the nibble is complete ----
        u32 f;
        if (k & 1u) {
            mbar_wait(P.mb_full1, ph_full1), ph_full1 ^= 1u;
            f = P.msg_full[1];
        } else {
            mbar_wait(P.mb_full0, ph_full0), ph_full0 ^= 1u;
            f = P.msg_full[0];
        }
        ++k;
        ZT(ts2 = clock64(); g_acc[8] += ts2 - ts1; g_acc[9] += ts2 - g_ts[1]);
#ifdef ZG_GUESS_STATS
        if (g_last < 16u) { ++g_tot; if (g_last == f) ++g_hit; } else ++g_none;
        g_last = 0xFFu;
#endif
        // ---- contexts of the next nibble ----
        const u32 win = f & 3u;
        if (c8base == 1u) {
            c8base = 16u | f;
        } else {
            const u32 c = ((c8base & 15u) << 4) | f;
            h = cx.next(c, pc);  // predictor.v:809-818
            cx.advance(c);
            c8base = 1u;
            byte_done = c;
        }
        const u32 next_vline = ((((h + 16u * c8base) * 16u) & (ht_len - 16u)) >> 6);
        // ---- the winning candidate's lanes act for the next nibble ----
        const bool nactor = powner && pcand == win;
        if (all_ok) {
            // the slot just used waits for the shadow of the next loads
            psl = sl, pslot_at = slot_at, pending = actor, pf = f, prev_vline = cur_vline;
            if (nactor) sl = qsl, slot_at = qat;
        } else {
            // some component's candidate could not be taken early: its slot is read now, after the
            // write-back of the slot in use (the reference updates the table in place)
            write_back(actor, sl, slot_at, f);
            __syncwarp();
            if (nactor) {
                if (q_ok) sl = qsl, slot_at = qat;
                else fresh(win, sl, slot_at);
            }
            ZT(g_ts[2] = clock64());
            pair_arrive(bar_id);
        }
        actor = nactor;
        cur_vline = next_vline;
        // ---- a nibble ahead: pull the lines the likeliest values of the next nibble lead to into L1 ----
        if (n_guess) {
            // The next slot's own bit histories say which way each node went more often (image in shared
            // memory: S published it).  Lanes 0..15 look at the nodes of the highest-order component,
            // lanes 16..31 at those of the ICM, which decides when the former has not seen this context.
            __syncwarp();
            const u32 gst = P.cand[(win * 8u + (lane < 16 ? u32(NI) : 0u)) * 16u + (u32(lane) & 15u)];
            const u32 lk = (likely[gst >> 5] >> (gst & 31u)) & 1u;
            const u32 ones = __ballot_sync(kFull, lk != 0u);
            const u32 seen = __ballot_sync(kFull, gst != 0u);
            const u32 mask = (seen & 2u) ? (ones & 0xFFFFu) : (ones >> 16);
            u32 node = 1;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                u32 bit = (mask >> node) & 1u;
                // variants: lane group 1 leaves the likeliest path at the last bit, 2 at the third, 3 at the second
                if (pcand != 0u && int(pcand) == 4 - b) bit ^= 1u;
                node = node * 2u + bit;
            }
#ifdef ZG_GUESS_STATS
            if (lane == 0) g_last = (seen & 0x20002u) ? (node & 15u) : 0xFFu;
#endif
            if (powner && pcand < n_guess) {
                const u32 g_key = key_of((c8base << 4) | (node & 15u));
                const u32 h0 = (g_key * 16u) & (ht_len - 16u);
                const u8 *b0 = slot_peek(h0);
                // not the line of the slot in use nor of the one whose write-back is still to come
                if (b0 && (h0 >> 6) != cur_vline && (h0 >> 6) != prev_vline) {
                    const u8 *line = reinterpret_cast<const u8 *>(reinterpret_cast<uintptr_t>(b0) & ~uintptr_t(63));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(line));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(line + 32));
                }
            }
        }
        ZT(g_acc[10] += clock64() - ts2);
    }
#ifdef ZG_GUESS_STATS
    if (lane == 0 && slot == 0) printf("guess: %u of %u right, %u without history\n", g_hit, g_tot, g_none);
#endif
#ifdef ZG_TIMING
    if (tslot0) {
        const double n = double(g_acc[3]);
        printf("per nibble (cycles): R wait at barrier %.0f (arrive->release %.0f), R nibble %.0f (restart->part %.0f, part->full %.0f)\n",
               g_acc[0] / n, g_acc[1] / n, g_acc[2] / n, g_acc[4] / n, g_acc[5] / n);
        printf("  S: part sent->awake %.0f, awake->candidates ready %.0f, ready->full awake %.0f (full sent->awake %.0f), tail %.0f, slow paths %.3f\n",
               g_acc[6] / n, g_acc[7] / n, g_acc[8] / n, g_acc[9] / n, g_acc[10] / n, g_acc[12] / n);
    }
#endif
}

// ------------------------------------------------------------------------------------------
// R: adaptive tables, rounds, arithmetic decoder, archive walk
// ------------------------------------------------------------------------------------------
struct RingIO {
    const u8 *ring;
    u64 pos;
    __device__ __forceinline__ u32 get() { return ring[(pos++) & (kRing - 1)]; }
};

__device__ __forceinline__ void ring_fill(u8 *ring, const u8 *base, u64 pos, u64 &filled, u64 limit, int lane) {
    if (filled < pos + 64) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u64 at = filled + u64(lane + 32 * q);
            ring[at & (kRing - 1)] = at < limit ? base[at] : u8(0);
        }
        filled += 128;
        __syncwarp();
    }
}

template <int NI>
struct Rounds {
    int2 *tabs;
    const u8 *cand;
    const int16_t *stretch;  // padded: entry 0 holds entry 1
    const u16 *squash;       // padded: indexed by p + 2048
    volatile u32 *msg_part, *msg_full;
    u32 mb_part, mb_full0, mb_full1;
    u32 node, yy, k;
    int depth, lane;
    bool tslot;
    int src[3];
};

// Four bits of one nibble through the tree; `win` selects the candidate image S published.
template <int NI>
__device__ __forceinline__ u32 decode_nibble(Rounds<NI> &T, u32 win, u32 &low, u32 &high, u32 &code, RingIO &io) {
    constexpr int NC = NI + 1;
    const u32 node = T.node;
    const int d = T.depth;
    const i32 t = T.yy ? 32767 : 0;
    // ---- this node's bit-history states and table entries (predictor.v:561, :622) ----
    const u8 *cs = T.cand + win * 128u + node;
    u32 st[NC];
    int2 e[NC], u[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) st[c] = cs[c * 16];
#pragma unroll
    for (int c = 0; c < NC; ++c) e[c] = T.tabs[c * 256 + st[c]];
    // where an ancestor's update lands on this node's entry: equal states, per component and level
    u32 pk[2] = {0, 0};
#pragma unroll
    for (int c = 0; c < NC; ++c) pk[c >> 2] |= st[c] << (8 * (c & 3));
    u32 eq[3][2];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const u32 anc = node >> max(d - l, 0);
#pragma unroll
        for (int w = 0; w < (NC + 3) / 4; ++w) {
            const u32 a = __shfl_sync(kFull, pk[w], int(anc));
            eq[l][w] = l < d ? __vcmpeq4(a, pk[w]) : 0u;
        }
    }
    u32 idx = 1;
    i32 sqf = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        // ---- predict this node from the current view of its entries (predictor.v:555-631) ----
        i32 p[NC];
        p[0] = e[0].y;
#pragma unroll
        for (int i = 1; i <= NI; ++i) p[i] = d_clamp2k((e[i].x * p[i - 1] + e[i].y * 64) >> 16);
        i32 sq[NC];
#pragma unroll
        for (int i = 1; i <= NI; ++i) sq[i] = T.squash[p[i] + 2048];
        sqf = NI > 0 ? sq[NI] : i32(T.squash[p[0] + 2048]);
        // ---- decode the bit of tree level l: its node finished predicting in this round ----
        const u32 p16 = u32(__shfl_sync(kFull, sqf, int(idx))) * 2u + 1u;
        const u32 mid = coder_mid(low, high, p16);
        const u32 y = code <= mid;
        if (y) high = mid; else low = mid + 1;
        // ---- the update outcome yy would cause (predictor.v:701-709, :776-791) ----
        {
            const u32 v0 = u32(e[0].x);
            const u32 v = u32(i32(v0) + ((t - i32(v0 >> 8)) >> 2));
            u[0] = make_int2(i32(v), i32(T.stretch[d_stretch_pad_idx(v >> 8)]));
        }
#pragma unroll
        for (int i = 1; i <= NI; ++i) {
            const i32 err = t - sq[i];
            u[i] = make_int2(d_clamp512k(e[i].x + ((err * p[i - 1] + 4096) >> 13)),
                             d_clamp512k(e[i].y + ((err + 16) >> 5)));
        }
        // ---- hand the update down the side of the tree that outcome leads to ----
        if (l < 3) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const i32 fx = __shfl_sync(kFull, u[c].x, T.src[l]);
                const i32 fy = __shfl_sync(kFull, u[c].y, T.src[l]);
                if ((eq[l][c >> 2] >> (8 * (c & 3))) & 1u) e[c] = make_int2(fx, fy);
            }
        }
        // ---- decoder renormalisation (decoder.v:104-117); low/high/code are warp-uniform ----
        while (__any_sync(kFull, (high ^ low) < 0x1000000u)) {
            low <<= 8;
            high = (high << 8) | 0xFFu;
            if (low == 0) low = 1;
            code = (code << 8) | io.get();
        }
        idx = idx * 2 + y;
        if (l == 1) {  // two bits known: S can request the four possible next slots
            if (T.lane == 0) {
                *T.msg_part = idx & 3u;
#ifdef ZG_TIMING
                if (T.tslot) { g_ts[0] = clock64(); g_acc[4] += g_ts[0] - g_ts[3]; }
#endif
                mbar_arrive(T.mb_part);
            }
        }
    }
    const u32 full = idx;  // 16 + the four bits
    if (T.lane == 0) {
        T.msg_full[T.k & 1u] = full & 15u;
#ifdef ZG_TIMING
        if (T.tslot) { g_ts[1] = clock64(); g_acc[5] += g_ts[1] - g_ts[0]; }
#endif
        mbar_arrive((T.k & 1u) ? T.mb_full1 : T.mb_full0);
    }
    ++T.k;
    // ---- the nodes on the decoded path learn (level order: a deeper node with the same state holds
    //      the later value) ----
    const bool mine = node != 0 && (full >> (4 - d)) == node && ((full >> (3 - d)) & 1u) == T.yy;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        if (mine && d == l) {
#pragma unroll
            for (int c = 0; c < NC; ++c) T.tabs[c * 256 + st[c]] = u[c];
        }
        __syncwarp();
    }
    return full & 15u;
}

template <int NI>
__device__ void run_rounds(const DecodeArgs &A, const PairMem &P, int bar_id, int slot, int bi, const u8 *smem) {
    constexpr int NC = NI + 1;
    const ModelDev &M = A.model;
    const int lane = threadIdx.x & 31;
    Rounds<NI> T;
    T.tabs = P.tabs, T.cand = P.cand;
    T.stretch = reinterpret_cast<const int16_t *>(smem);
    T.squash = reinterpret_cast<const u16 *>(smem + 65536);
    T.msg_part = P.msg_part, T.msg_full = P.msg_full;
    T.mb_part = P.mb_part, T.mb_full0 = P.mb_full0, T.mb_full1 = P.mb_full1;
    T.lane = lane, T.k = 0;
    T.tslot = slot == 0;
    T.node = u32(lane) & 15u, T.yy = u32(lane) >> 4;
    T.depth = 31 - __clz(int(T.node | 1u));
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        T.src[l] = lane;
        if (l < T.depth) {
            const u32 anc = T.node >> (T.depth - l), bit = (T.node >> (T.depth - l - 1)) & 1u;
            T.src[l] = int(anc + 16u * bit);
        }
    }
    {   // adaptive tables into shared memory: the fill kernel wrote their initial images into the workspace
        const u8 *ws = A.workspace + u64(slot) * M.ws_bytes;
        const u32 *src0 = reinterpret_cast<const u32 *>(ws + M.comps[0].cm_off);
        for (int q = lane; q < 256; q += 32) {
            const u32 v = src0[q];
            T.tabs[q] = make_int2(i32(v), i32(T.stretch[d_stretch_pad_idx(v >> 8)]));
        }
#pragma unroll
        for (int i = 1; i <= NI; ++i) {
            const int2 *src = reinterpret_cast<const int2 *>(ws + M.comps[i].cm_off);
            for (int q = lane; q < 256; q += 32) T.tabs[i * 256 + q] = src[q];
        }
        __syncwarp();
    }
    auto send = [&](u32 code_word) {
        if (lane == 0) {
            *P.msg_part = code_word;
            mbar_arrive(P.mb_part);
        }
    };
    const DecBlock blk = A.blocks[bi];
    const u8 *arc = A.arc;
    u64 pos = blk.arc_pos;  // uniform across the warp
    DecBlockOut res;
    res.end_pos = pos, res.out_len = 0, res.n_seg = 0, res.status = ZPAQGPU_OK;
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
        send(kSegBegin);
        // Decoder.init (decoder.v:29-47)
        u32 low = 1, high = 0xFFFFFFFFu, code = 0;
        u64 filled = pos;
        ring_fill(P.ring, arc, pos, filled, A.arc_len, lane);
        RingIO io{P.ring, pos};
        for (int q = 0; q < 4; ++q) code = (code << 8) | io.get();
        rec.out_off = blk.out_off + res.out_len;
        u64 produced = 0;
        int pp_state = 0;
        bool unsupported = false;
        u32 win = 0;
#ifdef ZG_TIMING
        const bool tslot0 = slot == 0 && lane == 0;
        long long tr0 = 0, tr1 = 0;
#endif
        for (;;) {
            ring_fill(P.ring, arc, io.pos, filled, A.arc_len, lane);
            // EOF flag: decode(p=0) => y = (code <= low) (decoder.v:128-131)
            const bool eof = code <= low;
            if (eof) high = low; else low = low + 1;
            while (__any_sync(kFull, (high ^ low) < 0x1000000u)) {
                low <<= 8;
                high = (high << 8) | 0xFFu;
                if (low == 0) low = 1;
                code = (code << 8) | io.get();
            }
            // the model is only consulted once a data byte is known to follow (decoder.v:128-142)
            if (eof) break;
            ZT(tr0 = clock64());
            pair_sync(bar_id);
            ZT(tr1 = clock64(); g_acc[0] += tr1 - tr0; g_acc[1] += tr1 - g_ts[2]; g_ts[3] = tr1);
            const u32 hi = decode_nibble<NI>(T, win, low, high, code, io);
            ZT(tr0 = clock64(); g_acc[2] += tr0 - tr1);
            pair_sync(bar_id);
            ZT(tr1 = clock64(); g_acc[0] += tr1 - tr0; g_acc[1] += tr1 - g_ts[2]; g_ts[3] = tr1);
            const u32 lo = decode_nibble<NI>(T, hi & 3u, low, high, code, io);
            ZT(tr0 = clock64(); g_acc[2] += tr0 - tr1; g_acc[3] += 2);
            win = lo & 3u;
            const u32 ch = (hi << 4) | lo;
            if (pp_state == 0) {  // PostProcessor.write state 0 (decompressor.v:58-70)
                pp_state = ch == 1 ? 2 : 1;
                if (pp_state == 2) { unsupported = true; break; }
            } else {
                ++produced;
            }
        }
        // S has announced slots for a nibble that is not decoded; take that arrival, then close the segment
        pair_sync(bar_id);
        send(kSegEnd);
        pair_sync(bar_id);
        if (unsupported) { res.status = ZPAQGPU_E_UNSUPPORTED; break; }
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
    send(kExit);
    if (pos > A.arc_len) pos = A.arc_len;
    res.end_pos = pos;
    if (lane == 0) A.results[bi] = res;
}

}  // namespace

// Warps 0..G-1 of a CTA are the R warps of its G blocks, warps G..2G-1 their S warps (so that the R
// warps spread over all four schedulers of the SM).
template <int NI>
__global__ void __launch_bounds__(448, 1) k_decode_tree2(DecodeArgs A, int pairs_per_cta) {
    extern __shared__ __align__(16) u8 smem[];
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(A.tables.stretch_pad);
        uint4 *d = reinterpret_cast<uint4 *>(smem);
        for (int q = threadIdx.x; q < 4096; q += blockDim.x) d[q] = g[q];
        const uint4 *g2 = reinterpret_cast<const uint4 *>(A.tables.squash_pad);
        uint4 *d2 = reinterpret_cast<uint4 *>(smem + 65536);
        for (int q = threadIdx.x; q < 512; q += blockDim.x) d2[q] = g2[q];
        u8 *s_nex = smem + 65536 + 8192;
        for (int q = threadIdx.x; q < 512; q += blockDim.x) s_nex[q] = A.tables.nex[q];
        if (threadIdx.x < 8) reinterpret_cast<u32 *>(smem + 65536 + 8192 + 512)[threadIdx.x] = A.tables.likely[threadIdx.x];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp >= pairs_per_cta ? 1 : 0;
    const int grp = role ? warp - pairs_per_cta : warp;
    const PairMem P = carve_pair<NI>(smem + kSharedTables + size_t(grp) * pair_smem_bytes(NI));
    if (role == 0 && lane == 0) {
        mbar_init(P.mb_part, 1), mbar_init(P.mb_full0, 1), mbar_init(P.mb_full1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int slot = blockIdx.x * pairs_per_cta + grp;
    if (slot >= A.n_blocks) return;
    const int bi = int(A.order[A.first_block + slot]);
    const int bar_id = 1 + grp;
    if (role == 0) run_rounds<NI>(A, P, bar_id, slot, bi, smem);
    else run_slots<NI>(A, P, bar_id, slot, bi, reinterpret_cast<const u16 *>(smem + 65536 + 8192),
                       reinterpret_cast<const u32 *>(smem + 65536 + 8192 + 512));
}

size_t tree2_smem_bytes(const Model &m, int pairs_per_cta) {
    return kSharedTables + size_t(pairs_per_cta) * pair_smem_bytes(m.n_isse);
}
int tree2_max_pairs_per_cta(const Model &m) {
    const size_t budget = 227 * 1024;
    int g = int((budget - kSharedTables) / pair_smem_bytes(m.n_isse));
    return g > 7 ? 7 : g;  // 14 warps; named barrier ids 1..7
}
bool tree2_supports(const Model &m) { return m.is_chain && !m.has_mix2 && m.n_isse <= 4; }

template <int NI>
static bool launch_t2(const DecodeArgs &D, int g, size_t smem, cudaStream_t s) {
    auto k = k_decode_tree2<NI>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) return false;
    const int grid = (D.n_blocks + g - 1) / g;
    k<<<grid, g * 64, smem, s>>>(D, g);
    return true;
}

bool launch_decode_tree2(const Model &m, const DecodeArgs &A, int pairs_per_cta, cudaStream_t s) {
    if (!tree2_supports(m)) return false;
    const size_t smem = tree2_smem_bytes(m, pairs_per_cta);
    switch (m.n_isse) {
    case 0: return launch_t2<0>(A, pairs_per_cta, smem, s);
    case 1: return launch_t2<1>(A, pairs_per_cta, smem, s);
    case 2: return launch_t2<2>(A, pairs_per_cta, smem, s);
    case 3: return launch_t2<3>(A, pairs_per_cta, smem, s);
    case 4: return launch_t2<4>(A, pairs_per_cta, smem, s);
    }
    return false;
}

}  // namespace zg

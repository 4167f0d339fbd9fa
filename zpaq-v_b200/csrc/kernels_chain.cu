// kernels_chain.cu -- the specialised hot-path kernel for models of the shape every predefined
// level has (levels.v:53-375):  ICM -> ISSE -> ... -> ISSE [-> MIX2 of the last two ISSEs], with
// the context hashes coming from one of the two HCOMP programs the levels use.
//
// Mapping (sm_100a): one ZPAQ block per warp, ONE MODEL COMPONENT PER LANE.
//  * lane 0 owns the ICM, lanes 1..NI the ISSEs.  Each lane keeps the 16-byte hash slot of its
//    component for the current nibble in four registers, extracts the bit-history state of the
//    current tree node with a shift, and reads its adaptive table entry (ICM {cm, stretch(cm>>8)},
//    ISSE {wt0, wt1}) from shared memory with one 64-bit load;
//  * the ISSE chain p[i] = clamp2k((wt0*p[i-1] + wt1*64) >> 16) is a serial dependency across
//    components: the 2*NI weights are all-gathered with warp shuffles and every lane evaluates the
//    chain redundantly in registers, so no lane waits on a neighbour's arithmetic;
//  * the arithmetic coder (low/high/code) is evaluated redundantly by all lanes out of registers:
//    the decoded bit is known warp-wide without a broadcast; code/plaintext bytes are staged
//    through shared-memory rings and moved to/from HBM by the whole warp;
//  * each lane updates its own table entry and inserts the successor state into its slot
//    registers; at a nibble boundary each lane writes its slot back to HBM with one 16-byte store
//    and probes the next one (Predictor.find_ht, predictor.v:495-532: three 16-byte loads inside
//    one 64-byte line);
//  * squash/stretch/next-state tables (72.5 KiB) are shared by the CTA in shared memory, each
//    block's adaptive tables (2 KiB per component) are private to its warp; MIX2 weights for the
//    256 possible c8 values of the current byte are staged in shared memory per byte.
#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"

#include <type_traits>

namespace zg {
namespace {

constexpr int kRing = 256;       // input ring (bytes), power of two
constexpr int kOutStage = 256;   // decoder plaintext stage (bytes)
constexpr unsigned kFull = 0xFFFFFFFFu;

__host__ __device__ constexpr size_t warp_smem_bytes(int ni, bool mix2) {
    return size_t(ni + 1) * 256 * 8        // ICM {cm, stretch} pairs, then ISSE weight pairs
           + 256                           // store sink for lanes that own no component
           + (mix2 ? 512 : 0)              // staged MIX2 weights
           + kRing + kOutStage;
}
constexpr size_t kSharedTables = 32768 * 2 + 4096 * 2 + 512;

__device__ __forceinline__ uint4 ldg128(const u8 *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// State and steps both chain decoders share: the closed-form context hashes, the staged MIX2
// weights, and the probing role.  For probing, lane = (component pc = lane & 7, candidate
// pcand = lane >> 3): the four lanes of a component hold its table geometry and context hash, request
// the four possible slot lines of the next nibble two bits early (probe_issue) and make the find_ht
// choice (choose) from registers when their guess was right.
template <int NI, bool MIX2>
struct ProbeBase {
    // shared-memory views
    const int16_t *stretch;  // padded: entry 0 holds entry 1
    const u16 *squash;       // padded: indexed by p + 2048
    const u16 *nex16;        // nex16[s] = next(s,0) | next(s,1) << 8
    u16 *a16s;
    u8 *ring;
    u8 *stage;
    int lane;
    // hash table of component pc (dense) or its page table (paged), and its context hash
    const ModelDev *md;
    u8 *ht;
    u32 ht_len;
    int sizebits;
    u32 h;
    int pc;
    u32 pcand;
    bool powner;             // pc <= NI
    // MIX2
    u16 *a16;
    u32 a16_mask, mix_h, mix_sel;
    i32 mix_rate;
    // context history (uniform)
    int ctx_mode, n_hash, n_comp;
    u32 hist;                // CTX_M1: previous three bytes; CTX_HASHCHAIN: previous byte
    // candidate slots of the NEXT nibble, requested two bits before its context is known
    uint4 q0, q1, q2;
    u8 *qb0;
    u32 q_key, cur_vline;    // cur_vline: 64-byte line (virtual offset >> 6) of the current slot
    bool q_ok, spec;

    __device__ void base_setup(const ModelDev &M, u8 *ws, const int16_t *st, const u16 *sq, const u8 *nx) {
        lane = threadIdx.x & 31;
        pc = lane & 7, pcand = u32(lane) >> 3;
        powner = pc <= NI;
        spec = true, q_ok = false, qb0 = nullptr, q_key = 0, cur_vline = ~0u;
        q0 = q1 = q2 = make_uint4(0, 0, 0, 0);
        stretch = st, squash = sq, nex16 = reinterpret_cast<const u16 *>(nx);
        ctx_mode = M.ctx_mode, n_hash = M.n_hash, n_comp = M.n;
        ht = nullptr, ht_len = 16, sizebits = 0;
        if (powner) {
            const CompDesc &cd = M.comps[pc];
            ht = ws + cd.ht_off, ht_len = cd.ht_len, sizebits = cd.a + 2;
        }
        md = &M;
        h = 0, hist = 0, mix_h = 0;
        a16 = nullptr, a16_mask = 0, mix_sel = 0, mix_rate = 0;
        if (MIX2) {
            const CompDesc &cd = M.comps[NI + 1];
            a16 = reinterpret_cast<u16 *>(ws + cd.a16_off);
            a16_mask = cd.a16_len - 1, mix_sel = cd.p[3], mix_rate = i32(cd.p[2]);
        }
    }

    // adaptive tables into shared memory: the fill kernel wrote their initial images into the workspace
    __device__ void load_tables(int2 *tables, const ModelDev &M, const u8 *ws) {
        const u32 *src0 = reinterpret_cast<const u32 *>(ws + M.comps[0].cm_off);
        for (int k = lane; k < 256; k += 32) {
            const u32 v = src0[k];
            tables[k] = make_int2(i32(v), i32(stretch[d_stretch_pad_idx(v >> 8)]));
        }
#pragma unroll
        for (int i = 1; i <= NI; ++i) {
            const int2 *src = reinterpret_cast<const int2 *>(ws + M.comps[i].cm_off);
            for (int k = lane; k < 256; k += 32) tables[i * 256 + k] = src[k];
        }
    }

    // pr.reset() (predictor.v:827-833): contexts go to zero, history and tables stay.
    __device__ void segment_reset() {
        h = 0, mix_h = 0;
        q_ok = false;  // requested with the old contexts
        stage_mix();
    }

    __device__ void stage_mix() {
        if (MIX2) {
            __syncwarp();
            for (int k = lane; k < 256; k += 32) a16s[k] = a16[(mix_h + u32(k)) & a16_mask];
            __syncwarp();
        }
    }

    // HCOMP in closed form (levels.v:72-87, :126-139): the context hash component `sel` gets for
    // the byte that follows byte c.
    __device__ __forceinline__ u32 ctx_next(u32 c, int sel, u32 &new_hist, u32 &mixv) const {
        u32 mine = 0;
        mixv = 0;
        if (ctx_mode == CTX_M1) {
            u32 a = (0u + c + 512u) * 773u;
            a = (a + (hist & 255u) + 512u) * 773u;
            const u32 h0 = a;
            a = (a + ((hist >> 8) & 255u) + 512u) * 773u;
            a = (a + ((hist >> 16) & 255u) + 512u) * 773u;
            mine = sel == 0 ? h0 : (sel == 1 ? a : 0u);
            new_hist = ((hist << 8) | c) & 0xFFFFFFu;
        } else {
            u32 a = c;
            for (int r = 0; r < n_hash; ++r) {
                a = (a + hist + 512u) * 773u;
                if (r == sel) mine = a;
                if (MIX2 && r == NI + 1) mixv = a;
            }
            new_hist = c;
        }
        return sel < n_comp ? mine : 0u;
    }

    // After byte c: latch the hash of component pc (predictor.v:809-818).
    __device__ void byte_end(u32 c) {
        u32 nh, mixv;
        h = ctx_next(c, pc, nh, mixv);
        hist = nh;
        if (MIX2) {
            mix_h = mixv;
            stage_mix();
        }
    }

    // Address of a slot without side effects: nullptr when a paged table has no page there yet.
    __device__ __forceinline__ u8 *slot_peek(u32 h0) const {
        if (!md->paged) return ht + h0;
        const u32 pte = reinterpret_cast<const u32 *>(ht)[h0 / kPageBytes];
        return pte ? md->pool + u64(pte - 1u) * kPageBytes + (h0 & (kPageBytes - 1u)) : nullptr;
    }

    // Called when the first two bits of a nibble are decoded (c8part = c8 after them).  The slot the
    // NEXT nibble probes is then one of four; lane (pc, pcand) requests the three candidate slots of
    // the line its completion pcand would lead to, so the HBM/L2 round trip runs under the rest of the
    // nibble, the table updates and the per-byte bookkeeping instead of after them.  Loads only: the
    // choice (and any eviction) happens in choose().  A line that is the current slot's own line is
    // not requested (it changes at the write-back), nor an unmapped page.
    __device__ __forceinline__ void probe_issue(u32 c8part) {
        q_ok = false;
        if (powner && spec) {
            const u32 c8new = (c8part << 2) | pcand;
            if (c8new < 256u) {
                q_key = h + 16u * c8new;            // low nibble of the same byte
            } else {
                u32 nh, mixv;                       // high nibble of the next byte (predictor.v:809-818)
                q_key = ctx_next(c8new & 255u, pc, nh, mixv) + 16u;
            }
            const u32 h0 = (q_key * 16u) & (ht_len - 16u);
            u8 *b0 = slot_peek(h0);
            if (b0 && (h0 >> 6) != cur_vline) {
                qb0 = b0;
                q0 = ldg128(b0);
                q1 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 16u));
                q2 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 32u));
                q_ok = true;
            }
        }
    }

    // Predictor.find_ht of component pc (predictor.v:495-532) for the nibble that starts with c8v.
    // Returns true on the acting lane of each component -- the lane whose early request was the right
    // one, else candidate lane 0, which loads now -- with the chosen slot and its address; `grp` has
    // bit 8k set when candidate lane k of this component acts.  Must run after the write-back of the
    // previous slot.
    __device__ __forceinline__ bool choose(u32 c8v, uint4 &chosen, u8 *&at, u32 &grp) {
        const u32 key = h + 16u * c8v;
        const bool match = powner && q_ok && q_key == key;
        grp = (__ballot_sync(kFull, match) >> pc) & 0x01010101u;
        const bool acting = powner && (grp ? match : pcand == 0u);
        const u32 h0 = (key * 16u) & (ht_len - 16u);
        cur_vline = h0 >> 6;
        if (acting) {
            const u32 chk = (key >> sizebits) & 255u;
            u8 *b0 = qb0;
            uint4 s0 = q0, s1 = q1, s2 = q2;
            if (!match) {
                // All three candidates are requested before any is looked at and the choice is made
                // with selects: as an if-chain the compiler serialises three HBM round trips.
                b0 = ht_slot(*md, ht, h0);
                asm volatile("" ::: "memory");  // after the write-back of the previous slot
                s0 = ldg128(b0);
                s1 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 16u));
                s2 = ldg128(reinterpret_cast<u8 *>(reinterpret_cast<uintptr_t>(b0) ^ 32u));
            }
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
        q_ok = false;
        return acting;
    }
};

// Serial decoder state: one model component per lane (lane i = component i, i <= NI).
template <int NI, bool MIX2>
struct Chain : ProbeBase<NI, MIX2> {
    using B = ProbeBase<NI, MIX2>;
    int2 *tab;         // this lane's table: ICM {cm[s], stretch(cm[s]>>8)} or ISSE {wt0, wt1}
    int2 *dump;        // 32 entries nobody reads
    u8 *slot_at;       // where the parked slot lives in HBM (nullptr: none yet)
    uint4 sl;          // the hash slot of component `lane` for the current nibble
    bool owner;        // lane <= NI

    __device__ void setup(u8 *smem_warp, const ModelDev &M, u8 *ws, const int16_t *st, const u16 *sq,
                          const u8 *nx) {
        B::base_setup(M, ws, st, sq, nx);
        owner = B::lane <= NI;
        u8 *p = smem_warp;
        int2 *tables = reinterpret_cast<int2 *>(p);
        p += size_t(NI + 1) * 2048;
        dump = reinterpret_cast<int2 *>(p), p += 256;
        B::a16s = reinterpret_cast<u16 *>(p), p += MIX2 ? 512 : 0;
        B::ring = p, p += kRing;
        B::stage = p;
        tab = tables + (owner ? B::lane : 0) * 256;
        B::load_tables(tables, M, ws);
        slot_at = nullptr;
        sl = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }

    // Predictor.find_ht for every component.  The slot of the previous nibble goes back to its table
    // first (the reference updates the table in place); the chosen slot then travels from the acting
    // lane to the component's own lane by SHFL.
    __device__ __forceinline__ void probe(u32 c8v) {
        if (owner && slot_at) *reinterpret_cast<uint4 *>(slot_at) = sl;
        uint4 nsl = make_uint4(0, 0, 0, 0);
        u8 *nat = nullptr;
        u32 grp;
        B::choose(c8v, nsl, nat, grp);
        // the acting lane of component pc: pc + 8 * (its candidate number), or pc itself
        const int from = grp ? B::pc + ((__ffs(int(grp)) - 1) & ~7) : B::pc;
        sl.x = __shfl_sync(kFull, nsl.x, from), sl.y = __shfl_sync(kFull, nsl.y, from);
        sl.z = __shfl_sync(kFull, nsl.z, from), sl.w = __shfl_sync(kFull, nsl.w, from);
        slot_at = reinterpret_cast<u8 *>(static_cast<uintptr_t>(
            __shfl_sync(kFull, u64(reinterpret_cast<uintptr_t>(nat)), from)));
    }
};

// Four coded bits of one nibble, whole warp converged (predictor.v:536-824 for the ICM/ISSE/MIX2
// chain, encoder.v:48-89 / decoder.v:73-118).  ENC: `nib` holds the 4 bits, MSB first; DEC: returns
// them.  c8 enters as 1 (high nibble) or 16..31 (low nibble).
template <int NI, bool MIX2, bool DEC, class IO>
__device__ __forceinline__ u32 code_nibble(Chain<NI, MIX2> &C, u32 nib, u32 &c8, u32 &low, u32 &high,
                                           u32 &code, IO &io) {
    const int lane = C.lane;
    u32 idx = 1;
    u32 got = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // ---- bit-history state of tree node idx in this lane's slot (predictor.v:561, :622) ----
        const u32 sh = (k == 0) ? 8u : (idx & 3u) * 8u;
        const bool hiword = (k == 3) && (idx & 4u);
        const u32 word = (k < 2) ? C.sl.x : (k == 2) ? C.sl.y : (hiword ? C.sl.w : C.sl.z);
        const u32 st = (word >> sh) & 255u;
        // ---- table reads of this bit ----
        const int2 e = C.tab[st];
        const u32 nx = C.nex16[st];
        i32 mw = 0;
        u32 msel = 0;
        if (MIX2) {
            msel = c8 & C.mix_sel;
            mw = C.a16s[msel];
        }
        // ICM update candidates for y=0 / y=1 and their stretch values (predictor.v:706-708);
        // meaningful on lane 0, harmless elsewhere
        const u32 v0 = u32(e.x);
        const i32 r0 = i32(v0 >> 8);
        const u32 va = u32(i32(v0) + ((0 - r0) >> 2)), vb = u32(i32(v0) + ((32767 - r0) >> 2));
        const i32 spa = C.stretch[d_stretch_pad_idx(va >> 8)], spb = C.stretch[d_stretch_pad_idx(vb >> 8)];
        // ---- predict: all-gather the weights, evaluate the ISSE chain on every lane ----
        i32 p = __shfl_sync(kFull, e.y, 0);  // p[0] = stretch(cm[state] >> 8)
        i32 pin = 0, pout = p, pa = 0, pb = 0;
#pragma unroll
        for (int i = 1; i <= NI; ++i) {
            const i32 w0 = __shfl_sync(kFull, e.x, i), w1 = __shfl_sync(kFull, e.y, i);
            const i32 pn = d_clamp2k((w0 * p + w1 * 64) >> 16);
            if (lane == i) pin = p, pout = pn;
            if (MIX2 && i == NI - 1) pa = pn;
            if (MIX2 && i == NI) pb = pn;
            p = pn;
        }
        if (MIX2) p = d_clamp2k((mw * pa + (65536 - mw) * pb) >> 16);
        const i32 sq_own = C.squash[pout + 2048];
        const i32 sq_fin = C.squash[p + 2048];
        const u32 p16 = u32(sq_fin) * 2u + 1u;
        // ---- code ----
        const u32 mid = coder_mid(low, high, p16);
        u32 y;
        if (DEC) {
            y = code <= mid;
        } else {
            y = (nib >> (3 - k)) & 1u;
        }
        if (y) high = mid; else low = mid + 1;
        while ((high ^ low) < 0x1000000u) {
            if (!DEC) io.put(high >> 24);
            low <<= 8;
            high = (high << 8) | 0xFFu;
            if (low == 0) low = 1;
            if (DEC) code = (code << 8) | io.get();
        }
        // ---- update this lane's component (predictor.v:701-709, :776-791) ----
        const i32 t = y ? 32767 : 0;
        const i32 err = t - sq_own;
        const i32 ix = d_clamp512k(e.x + ((err * pin + 4096) >> 13));
        const i32 iy = d_clamp512k(e.y + ((err + 16) >> 5));
        const i32 cx = i32(y ? vb : va), cy = y ? spb : spa;
        const bool icm = lane == 0;
        // non-owner lanes alias lane 0's table and write back what they read (st == 0 there is
        // never lane 0's live state only if values match, so they must not store at all)
        int2 *dstp = C.owner ? &C.tab[st] : &C.dump[lane];
        *dstp = make_int2(icm ? cx : ix, icm ? cy : iy);
        if (MIX2) {  // predictor.v:744-762, uniform values, one lane stores
            const i32 merr = ((t - sq_fin) * C.mix_rate) >> 5;
            i32 nw = mw + ((merr * (pa - pb) + 4096) >> 13);
            nw = max(0, min(65535, nw));
            if (lane == NI + 1) {
                C.a16s[msel] = u16(nw);
                C.a16[(C.mix_h + msel) & C.a16_mask] = u16(nw);
            }
        }
        // successor state back into the slot registers (statetable.v:75-88)
        const u32 ns = (nx >> (y * 8u)) & 255u;
        const u32 d = (st ^ ns) << sh;
        if (k < 2) C.sl.x ^= d;
        else if (k == 2) C.sl.y ^= d;
        else if (hiword) C.sl.w ^= d;
        else C.sl.z ^= d;
        c8 = (c8 << 1) | y;
        idx = (idx * 2 + y) & 15u;
        got = (got << 1) | y;
        if (DEC && k == 1) C.probe_issue(c8);
        if (MIX2) __syncwarp();  // the staged weight written by lane NI+1 may be re-read (mask < 255)
    }
    return got;
}

// ------------------------------------------------------------------------------------------
// Tree decoder: the whole nibble is predicted speculatively, one lane per (tree node, outcome).
//
// The decoder cannot know bit k+1's context before bit k is decoded, but inside one nibble the
// hash slot already holds the bit-history state of all 15 nodes of the nibble's binary context
// tree.  Lane L owns node n = L & 15 and assumes outcome yy = L >> 4 for it.  In four lock-step
// rounds (one per tree level) every lane evaluates the ICM/ISSE chain for its node from its own
// view of the adaptive table entries and prepares the entry update its outcome would cause; the
// update then reaches the descendants on that side of the node (one SHFL pair per component,
// applied only where the descendant's state equals the ancestor's -- the only way a bit of the
// nibble can influence a later one through the tables).  The arithmetic decoder then walks the
// tree with one SHFL per bit and the nodes on the decoded path store their updates.  The serial
// dependency per bit shrinks from "predict -> decode -> update -> next state -> predict" to the
// decoder step itself; the rounds run ahead of it.  Results are bit-identical to the serial
// evaluation (predictor.v:536-824): same table reads, same updates, in the same order.
// ------------------------------------------------------------------------------------------
template <int NI, bool MIX2>
struct Tree : ProbeBase<NI, MIX2> {
    using B = ProbeBase<NI, MIX2>;
    static constexpr int NC = NI + 1;
    int2 *tabs;        // NC tables of 256 entries: ICM {cm, stretch(cm>>8)}, ISSE {wt0, wt1}
    u8 *slots;         // NC x 16 bytes: the hash slots of the current nibble
    u8 *slot_at;       // acting lanes: where the current slot of component pc lives in HBM
    bool actor;        // this lane made the choice of the current nibble and writes the slot back
    // tree geometry of this lane
    u32 node, yy;
    int depth;
    int src[3];        // lane whose prepared update reaches this node after round l (self when none does)

    __device__ void setup(u8 *smem_warp, const ModelDev &M, u8 *ws, const int16_t *st, const u16 *sq,
                          const u8 *nx) {
        B::base_setup(M, ws, st, sq, nx);
        const int lane = B::lane;
        node = u32(lane) & 15u, yy = u32(lane) >> 4;
        depth = 31 - __clz(int(node | 1u));
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            src[l] = lane;
            if (l < depth) {
                const u32 anc = node >> (depth - l), bit = (node >> (depth - l - 1)) & 1u;
                src[l] = int(anc + 16u * bit);
            }
        }
        u8 *p = smem_warp;
        tabs = reinterpret_cast<int2 *>(p);
        p += size_t(NC) * 2048;
        slots = p, p += 256;
        B::a16s = reinterpret_cast<u16 *>(p), p += MIX2 ? 512 : 0;
        B::ring = p, p += kRing;
        B::stage = p;
        B::load_tables(tabs, M, ws);
        for (int k = lane; k < 64; k += 32) reinterpret_cast<u32 *>(slots)[k] = 0;
        slot_at = nullptr, actor = false;
        __syncwarp();
    }

    // Predictor.find_ht of every component; the acting lane publishes the chosen slot to the warp
    // through shared memory and keeps its address for the write-back at the end of the nibble.
    __device__ __forceinline__ void probe(u32 c8v) {
        uint4 sl = make_uint4(0, 0, 0, 0);
        u32 grp;
        actor = B::choose(c8v, sl, slot_at, grp);
        if (actor) *reinterpret_cast<uint4 *>(slots + 16 * B::pc) = sl;
        __syncwarp();
    }
};

// Four bits of one nibble through the tree (see above).  c8 enters as 1 (high nibble) or 16..31.
template <int NI, bool MIX2, class IO>
__device__ __forceinline__ void decode_nibble_tree(Tree<NI, MIX2> &T, u32 &c8, u32 &low, u32 &high, u32 &code,
                                                   IO &io) {
    constexpr int NC = NI + 1;
    const u32 node = T.node;
    const int d = T.depth;
    const i32 t = T.yy ? 32767 : 0;
    // ---- this node's bit-history states and table entries (predictor.v:561, :622) ----
    u32 st[NC], nx[NC];
    int2 e[NC], u[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) st[c] = T.slots[c * 16 + node];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        e[c] = T.tabs[c * 256 + st[c]];
        nx[c] = T.nex16[st[c]];
    }
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
    i32 mw = 0;
    u32 msel = 0;
    i32 nw = 0;
    if (MIX2) {
        msel = ((c8 << d) | (node - (1u << d))) & T.mix_sel;
        mw = T.a16s[msel];
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
        if (MIX2) {
            const i32 pf = d_clamp2k((mw * p[NI - 1] + (65536 - mw) * p[NI]) >> 16);
            sqf = T.squash[pf + 2048];
        } else {
            sqf = NI > 0 ? sq[NI] : i32(T.squash[p[0] + 2048]);
        }
        // ---- decode the bit of tree level l: its node finished predicting in this round ----
        const u32 p16 = u32(__shfl_sync(kFull, sqf, int(idx))) * 2u + 1u;
        const u32 mid = coder_mid(low, high, p16);
        const u32 y = code <= mid;
        if (y) high = mid; else low = mid + 1;
        // ---- the update outcome yy would cause (predictor.v:701-709, :776-791, :744-762) ----
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
        if (MIX2) {
            const i32 merr = ((t - sqf) * T.mix_rate) >> 5;
            nw = max(0, min(65535, mw + ((merr * (p[NI - 1] - p[NI]) + 4096) >> 13)));
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
        // ---- decoder renormalisation (decoder.v:104-117) ----
        while ((high ^ low) < 0x1000000u) {
            low <<= 8;
            high = (high << 8) | 0xFFu;
            if (low == 0) low = 1;
            code = (code << 8) | io.get();
        }
        idx = idx * 2 + y;
        if (l == 1) T.probe_issue((c8 << 2) | (idx & 3u));
    }
    const u32 full = idx;  // 16 + the four bits
    // ---- the nodes on the decoded path learn: table entries, successor states ----
    const bool mine = node != 0 && (full >> (4 - d)) == node && ((full >> (3 - d)) & 1u) == T.yy;
#pragma unroll
    for (int l = 0; l < 4; ++l) {  // level order: a deeper node with the same state holds the later value
        if (mine && d == l) {
#pragma unroll
            for (int c = 0; c < NC; ++c) T.tabs[c * 256 + st[c]] = u[c];
        }
        __syncwarp();
    }
    if (mine) {
#pragma unroll
        for (int c = 0; c < NC; ++c) T.slots[c * 16 + node] = u8((nx[c] >> (T.yy * 8u)) & 255u);
        if (MIX2) {
            T.a16s[msel] = u16(nw);
            T.a16[(T.mix_h + msel) & T.a16_mask] = u16(nw);
        }
    }
    __syncwarp();
    if (T.actor) *reinterpret_cast<uint4 *>(T.slot_at) = *reinterpret_cast<const uint4 *>(T.slots + 16 * T.pc);
    c8 = (c8 << 4) | (full & 15u);
}

// ---- byte I/O over the shared-memory stages (all lanes execute, same address) ----
struct DecIO {
    const u8 *ring;
    u64 pos;
    __device__ __forceinline__ void put(u32) {}
    __device__ __forceinline__ u32 get() { return ring[(pos++) & (kRing - 1)]; }
};

// Keep at least 64 bytes of [pos, limit) in the ring; bytes past `limit` read as zero (the
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

__device__ __forceinline__ void load_shared_tables(u8 *smem, const DevTables &T) {
    const uint4 *g = reinterpret_cast<const uint4 *>(T.stretch_pad);
    uint4 *d = reinterpret_cast<uint4 *>(smem);
    for (int k = threadIdx.x; k < 4096; k += blockDim.x) d[k] = g[k];
    const uint4 *g2 = reinterpret_cast<const uint4 *>(T.squash_pad);
    uint4 *d2 = reinterpret_cast<uint4 *>(smem + 65536);
    for (int k = threadIdx.x; k < 512; k += blockDim.x) d2[k] = g2[k];
    u8 *s_nex = smem + 65536 + 8192;
    for (int k = threadIdx.x; k < 512; k += blockDim.x) s_nex[k] = T.nex[k];
    __syncthreads();
}

}  // namespace

// ------------------------------------------------------------------------------------------
// k_decode_chain
// ------------------------------------------------------------------------------------------
template <int NI, bool MIX2, bool TREE>
__global__ void __launch_bounds__(256, 1) k_decode_chain(DecodeArgs A) {
    extern __shared__ __align__(16) u8 smem[];
    load_shared_tables(smem, A.tables);
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + wic;
    if (slot >= A.n_blocks) return;
    const int bi = int(A.order[A.first_block + slot]);
    typename std::conditional<TREE, Tree<NI, MIX2>, Chain<NI, MIX2>>::type C;
    u8 *ws = A.workspace + u64(slot) * A.model.ws_bytes;
    C.setup(smem + kSharedTables + size_t(wic) * warp_smem_bytes(NI, MIX2), A.model, ws,
            reinterpret_cast<const int16_t *>(smem), reinterpret_cast<const u16 *>(smem + 65536),
            smem + 65536 + 8192);
    C.spec = (A.flags & 1) != 0;
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
            // EOF flag: decode(p=0) => y = (code <= low) (decoder.v:128-131)
            const bool eof = code <= low;
            if (eof) high = low; else low = low + 1;
            while ((high ^ low) < 0x1000000u) {
                low <<= 8;
                high = (high << 8) | 0xFFu;
                if (low == 0) low = 1;
                code = (code << 8) | io.get();
            }
            if (eof) break;
            // the model is only consulted once a data byte is known to follow (decoder.v:128-142):
            // a probe at EOF could evict a slot that a later segment of the block still needs
            u32 c8 = 1;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                C.probe(c8);
                if constexpr (TREE) decode_nibble_tree<NI, MIX2>(C, c8, low, high, code, io);
                else code_nibble<NI, MIX2, true>(C, 0, c8, low, high, code, io);
            }
            const u32 ch = c8 & 255u;
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

template <int NI, bool MIX2, bool TREE>
static bool launch_dec(const DecodeArgs &D, int wpc, size_t smem, cudaStream_t s) {
    const int grid = (D.n_blocks + wpc - 1) / wpc;
    auto k = k_decode_chain<NI, MIX2, TREE>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) return false;
    k<<<grid, wpc * 32, smem, s>>>(D);
    return true;
}

bool launch_decode_chain(const Model &m, const DecodeArgs &A, int wpc, bool tree, cudaStream_t s) {
    if (!m.is_chain) return false;
    const size_t smem = chain_smem_bytes(m, wpc);
    // The tree decoder evaluates all components of a node in one lane, the serial decoder spreads the
    // components over lanes: measured on B200 (512 x 128 KiB text) the tree wins at -m1/-m2/-m3
    // (-16/-15/-11 % kernel time) and loses at -m4/-m5 (six and eight components plus MIX2, +33/+30 %).
    // It also takes MIX2 weights per node from the staged copy, which needs mask 255.
    if (m.has_mix2 || m.n_isse > 4) tree = false;
#define ZG_CASE(NI, MX)                                                        \
    if (m.n_isse == NI && m.has_mix2 == MX)                                    \
        return tree ? launch_dec<NI, MX, true>(A, wpc, smem, s) : launch_dec<NI, MX, false>(A, wpc, smem, s);
    ZG_CASE(0, false) ZG_CASE(1, false) ZG_CASE(2, false) ZG_CASE(3, false) ZG_CASE(4, false)
    ZG_CASE(5, false) ZG_CASE(6, false) ZG_CASE(7, false)
    ZG_CASE(2, true) ZG_CASE(3, true) ZG_CASE(4, true) ZG_CASE(5, true) ZG_CASE(6, true) ZG_CASE(7, true)
#undef ZG_CASE
    return false;
}

}  // namespace zg

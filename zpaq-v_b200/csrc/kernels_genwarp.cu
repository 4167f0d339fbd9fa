// kernels_genwarp.cu -- the all-components codec kernel as a warp program: any header the reference
// decoder accepts (decompressor.v:278-342) with up to 32 components -- CONS/CM/ICM/MATCH/AVG/MIX2/MIX/
// ISSE/SSE in any wiring -- and an arbitrary HCOMP program.
//
// Mapping (sm_100a), one ZPAQ block per warp:
//  * COMPONENT i LIVES ON LANE i: its Component fields (a, b, c, limit, cxt), its context hash h[i], its
//    prediction p[i], the 16-byte hash slot of the current nibble (ICM/ISSE: loaded at the nibble boundary,
//    bit-history states read and replaced in registers, written back at the next boundary) and its table
//    pointers sit in that lane's registers.  The small adaptive tables (ICM cm[256], ISSE cm[512]) are
//    copied into the warp's shared memory while they fit its budget; the large ones (hash tables, CM, MIX,
//    MIX2, SSE, MATCH) stay in the block's HBM workspace.
//  * predict (predictor.v:536-668) runs in two parts.  FETCH: every lane does what does not depend on
//    another component -- the hash-slot search of ICM/ISSE (find_ht, three candidates of one 64-byte line),
//    the bit-history state, the table entry, the MIX2 weight, CONS/CM/ICM/MATCH predictions.  Lanes of one
//    type run together; different types take different branches one after the other.  COMBINE: components
//    that read other predictions (AVG/MIX2/MIX/ISSE/SSE) are evaluated level by level (level = 1 + deepest
//    input in front of it, from the host); inputs travel by SHFL.  The reference evaluates in index order
//    and lets a component read p[j] of a LATER component (it only checks j < n): that value is the one of
//    the previous bit, kept here as `pp`.  A MIX is a warp dot product: lane j+l multiplies its own
//    prediction with weight l, the sum is one __reduce_add_sync (wrapping 32-bit adds, order-free); its
//    update is one weight per lane.
//    Measured (ncu, profiles/r02_genwarp_decode_lines.txt): the kernel is bound by its instruction stream
//    (3 300 warp instructions per nibble for a five-type header, 5 % waiting for memory), so what the lanes
//    buy is one pass for all components of a type; issuing the loads of all types from one place first was
//    tried and was slower (743 against 669 ms: more instructions, nothing to hide).
//  * update (predictor.v:672-824) touches only a component's own tables and reads the finished p[]: all
//    lanes at once.
//  * the ZPAQL interpreter (zpaql.v:167-954) is WARP-UNIFORM: a, b, c, d, f, pc are replicated in every
//    lane, every lane decodes the same instruction, loads from M/H/R are same-address loads, stores are
//    done by lane 0.  After the run lane i picks up h[i] itself.  R, and H and M when they fit, live in
//    the warp's shared memory, else in the workspace.
//  * the arithmetic coder runs redundantly in all lanes (no broadcast of the decoded bit); squash,
//    stretch and the next-state table are shared by the CTA in shared memory.
#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"

namespace zg {
namespace {

constexpr unsigned kAll = 0xFFFFFFFFu;
constexpr int kWarpsPerCta = 4;
constexpr size_t kConstBytes = 32768 * 2 + 4096 * 2 + 512;  // stretch_pad, squash_pad, next-state pairs
constexpr size_t kVmBytes = 4096;                            // per warp: R (1 KiB) + H and M when they fit
constexpr size_t kTabBytes = 24576;                          // per warp: ICM / ISSE adaptive tables
constexpr size_t kWarpBytes = kVmBytes + kTabBytes;

__device__ __forceinline__ uint4 ldg128(const u8 *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------
// ZPAQL, warp-uniform
// ------------------------------------------------------------------------------------------
struct VmU {
    u32 a, b, c, d;
    i32 f, pc;
    u8 *m;
    u32 *h;
    u32 *r;
    u32 m_mask, h_mask;
    bool has_m, has_h;
    const u8 *hdr;
    i32 hbegin, hend, hdr_len;
    int lane;

    __device__ __forceinline__ u32 mget(u32 i) const { return has_m ? m[i & m_mask] : 0u; }
    __device__ __forceinline__ void mset(u32 i, u32 v) {
        if (has_m) {
            if (lane == 0) m[i & m_mask] = u8(v);
            __syncwarp();
        }
    }
    __device__ __forceinline__ u32 hget(u32 i) const { return has_h ? h[i & h_mask] : 0u; }
    __device__ __forceinline__ void hset(u32 i, u32 v) {
        if (has_h) {
            if (lane == 0) h[i & h_mask] = v;
            __syncwarp();
        }
    }
    __device__ __forceinline__ u32 rget(i32 i) const { return r[i & 255]; }
    __device__ __forceinline__ void rset(i32 i, u32 v) {
        if (lane == 0) r[i & 255] = v;
        __syncwarp();
    }
    __device__ u32 src(int y, i32 operand) const {
        switch (y) {
        case 0: return a;
        case 1: return b;
        case 2: return c;
        case 3: return d;
        case 4: return mget(b);
        case 5: return mget(c);
        case 6: return hget(d);
        default: return u32(operand);
        }
    }
    __device__ __forceinline__ void set_reg(int x, u32 v) {
        switch (x) {
        case 0: a = v; break;
        case 1: b = v; break;
        case 2: c = v; break;
        default: d = v; break;
        }
    }
    __device__ __forceinline__ u32 get_reg(int x) const { return x == 0 ? a : x == 1 ? b : x == 2 ? c : d; }

    // One instruction, every lane the same one; false stops the run (zpaql.v:215-954).
    __device__ bool step() {
        if (pc < hbegin || pc >= hend) return false;
        const u32 op = hdr[pc++];
        i32 operand = 0;
        if (op == 255) {  // 3-byte form, types.v:51-64
            if (pc + 1 < hdr_len) {
                operand = i32(hdr[pc]) + i32(hdr[pc + 1]) * 256;
                pc += 2;
            }
        } else if ((op & 7) == 7) {
            if (pc < hdr_len) operand = hdr[pc++];
        }
        if (op >= 64 && op < 120) {  // X=Y
            const u32 v = src(op & 7, operand);
            const int x = (op >> 3) & 7;
            if (x < 4) set_reg(x, v);
            else if (x == 4) mset(b, v);
            else if (x == 5) mset(c, v);
            else hset(d, v);
            return true;
        }
        if (op >= 128 && op < 240) {  // a op= Y, comparisons set f
            const u32 v = src(op & 7, operand);
            switch ((op - 128) >> 3) {
            case 0: a += v; break;
            case 1: a -= v; break;
            case 2: a *= v; break;
            case 3: if (v) a /= v; break;   // x/0 keeps a (zpaql.v:697-741)
            case 4: if (v) a %= v; break;
            case 5: a &= v; break;
            case 6: a &= ~v; break;
            case 7: a |= v; break;
            case 8: a ^= v; break;
            case 9: a <<= (v & 31); break;
            case 10: a >>= (v & 31); break;
            case 11: f = (a == v); break;
            case 12: f = (a < v); break;
            default: f = (a > v); break;
            }
            return true;
        }
        if (op < 32 && (op & 7) < 5) {  // swap / ++ / -- / ~ / =0 on a, b, c, d
            const int x = op >> 3;
            const u32 v = get_reg(x);
            switch (op & 7) {
            case 0:
                if (op) {  // op 0 is NOP; 8/16/24 swap with a
                    set_reg(x, a);
                    a = v;
                }
                break;
            case 1: set_reg(x, v + 1); break;
            case 2: set_reg(x, v - 1); break;
            case 3: set_reg(x, ~v); break;
            default: set_reg(x, 0); break;
            }
            return true;
        }
        switch (op) {
        case 7: a = rget(operand); break;
        case 15: b = rget(operand); break;
        case 23: c = rget(operand); break;
        case 31: d = rget(operand); break;
        case 32: { const u32 t = mget(b); mset(b, a); a = t; break; }
        case 33: mset(b, mget(b) + 1); break;
        case 34: mset(b, mget(b) - 1); break;
        case 35: mset(b, ~mget(b)); break;
        case 36: mset(b, 0); break;
        case 39: if (f != 0) pc += ((operand + 128) & 255) - 127; break;
        case 40: { const u32 t = mget(c); mset(c, a); a = t; break; }
        case 41: mset(c, mget(c) + 1); break;
        case 42: mset(c, mget(c) - 1); break;
        case 43: mset(c, ~mget(c)); break;
        case 44: mset(c, 0); break;
        case 47: if (f == 0) pc += ((operand + 128) & 255) - 127; break;
        case 48: { const u32 t = hget(d); hset(d, a); a = t; break; }
        case 49: hset(d, hget(d) + 1); break;
        case 50: hset(d, hget(d) - 1); break;
        case 51: hset(d, ~hget(d)); break;
        case 52: hset(d, 0); break;
        case 55: rset(operand, a); break;
        case 56: return false;  // HALT
        case 57: break;         // OUT has no observer on the HCOMP path
        case 59: a = (a + mget(b) + 512u) * 773u; break;
        case 60: hset(d, (hget(d) + a + 512u) * 773u); break;
        case 63: pc += ((operand + 128) & 255) - 127; break;
        case 255:
            pc = hbegin + i32(hdr[pc - 2]) + i32(hdr[pc - 1]) * 256;
            if (pc >= hend) return false;
            break;
        default: return false;
        }
        return true;
    }

    __device__ void run(u32 input) {
        a = input;
        pc = hbegin;
        while (pc < hend && pc >= hbegin)
            if (!step()) break;
    }
};

// ------------------------------------------------------------------------------------------
// The predictor: one component per lane
// ------------------------------------------------------------------------------------------
struct GenW {
    // ---- this lane's component ----
    bool act;              // lane < n
    i32 type, a, b, c, limit;
    u32 cxt, h;
    i32 p, pp;             // prediction of this bit / of the previous bit
    i32 lvl;
    u32 *cm;
    u8 *ht;
    u16 *a16;
    u32 cm_len, ht_len, q0, q1, q2, q3;
    i32 ja, jb;            // the lanes this component reads (AVG j,k; MIX2 j,k; ISSE/SSE j)
    bool va, vb;           // ... and whether they name a component (j < n)
    i32 e0, e1;            // fetched table entry (ISSE weights, MIX2 weight, CM / SSE cells)
    uint4 sl;              // ICM / ISSE: the hash slot of the current nibble
    u8 *slot_at;           // ... and where it lives in the table (nullptr: none yet)
    bool sse_ok, sse_hi;   // SSE: e0/e1 hold cm[idx], cm[idx+1] of this bit; the cell update will touch is the second
    // ---- uniform ----
    i32 n, n_levels, lane;
    u32 c8, hmap4, h_len;
    VmU vm;
    const int16_t *stretch_pad;
    const u16 *squash_pad;
    const u8 *nex;
    const i32 *dt, *dt2k;

    __device__ __forceinline__ i32 squash(i32 d) const { return squash_pad[max(0, min(4095, d + 2048))]; }
    __device__ __forceinline__ i32 stretch(i32 x) const { return stretch_pad[max(0, min(32767, x))]; }

    __device__ void block_init(const ModelDev &M, u8 *ws, const DevTables &T, const u8 *smem_const, u8 *smem_vm) {
        lane = threadIdx.x & 31;
        n = M.n;
        act = lane < n;
        stretch_pad = reinterpret_cast<const int16_t *>(smem_const);
        squash_pad = reinterpret_cast<const u16 *>(smem_const + 65536);
        nex = smem_const + 65536 + 8192;
        dt = T.dt, dt2k = T.dt2k;
        type = C_NONE, a = b = c = limit = 0, cxt = 0, h = 0, p = pp = 0, lvl = 0;
        cm = nullptr, ht = nullptr, a16 = nullptr, cm_len = ht_len = 1, q0 = q1 = q2 = q3 = 0;
        ja = jb = 0, va = vb = false, e0 = e1 = 0;
        if (act) {
            const CompDesc &cd = M.comps[lane];
            type = cd.type, a = cd.a, b = cd.b, c = cd.c, limit = cd.limit, lvl = cd.level;
            cm = reinterpret_cast<u32 *>(ws + cd.cm_off), ht = ws + cd.ht_off;
            a16 = reinterpret_cast<u16 *>(ws + cd.a16_off);
            cm_len = cd.cm_len, ht_len = cd.ht_len;
            q0 = cd.p[0], q1 = cd.p[1], q2 = cd.p[2], q3 = cd.p[3];
            i32 j = 0, k = 0;
            switch (type) {
            case C_AVG: j = a, k = b; break;
            case C_MIX2: j = i32(q0), k = i32(q1); break;
            case C_ISSE: case C_SSE: j = k = b; break;
            default: break;
            }
            va = j < n, vb = k < n;
            ja = va ? j : 0, jb = vb ? k : 0;
        }
        sl = make_uint4(0, 0, 0, 0), slot_at = nullptr, sse_ok = sse_hi = false;
        n_levels = __reduce_max_sync(kAll, act ? lvl : 0) + 1;
        {   // ICM cm[256] / ISSE cm[512] into the warp's shared memory while the budget lasts (index order)
            const u32 need = (act && type == C_ICM) ? 1024u : (act && type == C_ISSE) ? 2048u : 0u;
            u32 incl = need;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 up = __shfl_up_sync(kAll, incl, d);
                if (lane >= d) incl += up;
            }
            const u32 at = incl - need;
            const bool in_smem = need != 0 && at + need <= u32(kTabBytes);
            u32 *dst = reinterpret_cast<u32 *>(smem_vm + kVmBytes + at);
            for (int i = 0; i < n; ++i) {
                const u32 words = __shfl_sync(kAll, in_smem ? need / 4 : 0u, i);
                if (!words) continue;
                const u64 from = __shfl_sync(kAll, u64(reinterpret_cast<uintptr_t>(cm)), i);
                const u64 to = __shfl_sync(kAll, u64(reinterpret_cast<uintptr_t>(dst)), i);
                const u32 *srcp = reinterpret_cast<const u32 *>(static_cast<uintptr_t>(from));
                u32 *dstp = reinterpret_cast<u32 *>(static_cast<uintptr_t>(to));
                for (u32 k = lane; k < words; k += 32) dstp[k] = srcp[k];
            }
            __syncwarp();
            if (in_smem) cm = dst;
        }
        // ZPAQL state (zpaql.v:74-96); cleared by the workspace memset / here for the shared part
        h_len = M.h_len;
        vm.lane = lane;
        vm.a = vm.b = vm.c = vm.d = 0, vm.f = 0, vm.pc = M.hbegin;
        vm.has_m = M.m_len != 0, vm.has_h = M.h_len != 0;
        vm.m_mask = M.m_len - 1, vm.h_mask = M.h_len - 1;
        vm.hdr = M.header, vm.hbegin = M.hbegin, vm.hend = M.hend, vm.hdr_len = M.header_len;
        vm.r = reinterpret_cast<u32 *>(smem_vm);
        const size_t hm_bytes = size_t(M.h_len) * 4 + M.m_len;
        if (hm_bytes <= kVmBytes - 1024) {
            vm.h = reinterpret_cast<u32 *>(smem_vm + 1024);
            vm.m = smem_vm + 1024 + size_t(M.h_len) * 4;
        } else {
            vm.h = reinterpret_cast<u32 *>(ws + M.h_off);
            vm.m = ws + M.m_off;
        }
        for (int k = lane; k < int(kVmBytes / 4); k += 32) reinterpret_cast<u32 *>(smem_vm)[k] = 0;
        __syncwarp();
        c8 = 1, hmap4 = 1;
    }

    __device__ void segment_reset() {  // predictor.v:827-833
        c8 = 1, hmap4 = 1;
        h = 0;
    }

    // Predictor.find_ht (predictor.v:495-532) at a nibble boundary.  The slot of the previous nibble goes back
    // to its table first (the reference updates the table in place; nobody but this lane reads this table),
    // the three candidates of the line are requested together and the choice is made from registers.  An
    // evicted slot is cleared in the registers; the table sees it at the next write-back.
    __device__ __forceinline__ void probe(i32 sizebits, u32 key) {
        if (slot_at) *reinterpret_cast<uint4 *>(slot_at) = sl;
        const u32 chk = (key >> sizebits) & 255u;
        const u32 h0 = (key * 16u) & (ht_len - 16u);
        u8 *b0 = ht + h0, *b1 = ht + (h0 ^ 16u), *b2 = ht + (h0 ^ 32u);
        const uint4 s0 = ldg128(b0), s1 = ldg128(b1), s2 = ldg128(b2);
        const bool m0 = (s0.x & 255u) == chk, m1 = (s1.x & 255u) == chk, m2 = (s2.x & 255u) == chk;
        const u32 p0 = (s0.x >> 8) & 255u, p1 = (s1.x >> 8) & 255u, p2 = (s2.x >> 8) & 255u;
        u8 *victim = (p0 <= p1 && p0 <= p2) ? b0 : (p1 < p2 ? b1 : b2);
        const bool hit = m0 | m1 | m2;
        slot_at = m0 ? b0 : m1 ? b1 : m2 ? b2 : victim;
        const uint4 pick = m0 ? s0 : (m1 ? s1 : s2);
        sl.x = hit ? pick.x : chk, sl.y = hit ? pick.y : 0u, sl.z = hit ? pick.z : 0u, sl.w = hit ? pick.w : 0u;
    }
    // byte i (1..15) of the slot registers: the bit-history state of tree node i (predictor.v:561, :622)
    __device__ __forceinline__ u32 slot_state(u32 i) const {
        const u32 w = i < 8 ? (i < 4 ? sl.x : sl.y) : (i < 12 ? sl.z : sl.w);
        return (w >> ((i & 3u) * 8u)) & 255u;
    }
    __device__ __forceinline__ void slot_set(u32 i, u32 old_state, u32 new_state) {
        const u32 d = (old_state ^ new_state) << ((i & 3u) * 8u);
        if (i < 4) sl.x ^= d;
        else if (i < 8) sl.y ^= d;
        else if (i < 12) sl.z ^= d;
        else sl.w ^= d;
    }

    // p[j] as component `lane` sees it during predict: this bit's value when j is in front of it, else
    // the previous bit's (the reference reads rt[j].p before component j has been evaluated)
    __device__ __forceinline__ i32 input(i32 j) const {
        const i32 cur = __shfl_sync(kAll, p, j), old = __shfl_sync(kAll, pp, j);
        return j < lane ? cur : old;
    }

    __device__ i32 predict() {  // predictor.v:536-668; returns squash(p[n-1])
        pp = p;
        // ---- FETCH: everything that needs no other component ----
        if (act) {
            const bool nibble = c8 == 1 || (c8 & 0xf0) == 16;
            switch (type) {
            case C_CONS: p = (a - 128) * 16; break;
            case C_CM:
                cxt = h ^ hmap4;
                e0 = i32(cm[cxt & (cm_len - 1)]);
                p = stretch(i32(u32(e0) >> 17));
                break;
            case C_ICM:
                if (nibble) probe(a + 2, h + 16u * c8);
                cxt = slot_state(hmap4 & 15u);
                p = stretch(i32(cm[cxt] >> 8));
                break;
            case C_MATCH:
                if (a == 0) {
                    p = 0;
                } else {
                    const i32 idx = (limit - b) & i32(ht_len - 1);
                    c = i32((u32(ht[idx]) >> (7 - i32(cxt))) & 1u);
                    p = stretch((dt2k[a & 255] * (c * -2 + 1)) & 32767);
                }
                break;
            case C_MIX2:
                cxt = (h + (c8 & q3)) & u32(c - 1);
                e0 = a16[cxt];
                break;
            case C_MIX:
                cxt = u32((i32(h) + (i32(c8) & i32(q1))) & (c - 1));
                break;
            case C_ISSE: {
                if (nibble) probe(a + 2, h + 16u * c8);
                cxt = slot_state(hmap4 & 15u);
                const uint2 w = *reinterpret_cast<const uint2 *>(cm + cxt * 2);
                e0 = i32(w.x), e1 = i32(w.y);
                break;
            }
            case C_SSE: cxt = (h + c8) * 32u; break;
            case C_AVG: break;
            default: p = 0; break;
            }
        }
        // ---- COMBINE: level by level ----
        for (i32 lv = 1; lv < n_levels; ++lv) {
            const i32 pj = input(ja), pk = input(jb);
            const bool now = act && lvl == lv;
            if (now) {
                switch (type) {
                case C_AVG: p = (va && vb) ? ((pj * c + pk * (256 - c)) >> 8) : 0; break;
                case C_MIX2: p = (va && vb) ? d_clamp2k((e0 * pj + (65536 - e0) * pk) >> 16) : 0; break;
                case C_ISSE: p = va ? d_clamp2k((e0 * pj + e1 * 64) >> 16) : d_clamp2k(e1 >> 10); break;
                case C_SSE: {
                    i32 pq = va ? pj + 992 : 992;
                    pq = max(0, min(1983, pq));
                    const i32 wt = pq & 63;
                    pq >>= 6;
                    const i32 idx = i32(cxt) + pq;
                    sse_ok = idx >= 0 && idx + 1 < i32(cm_len);
                    if (sse_ok) {
                        e0 = i32(cm[idx]), e1 = i32(cm[idx + 1]);
                        const i32 p1 = i32(u32(e0) >> 10), p2 = i32(u32(e1) >> 10);
                        p = stretch((p1 * (64 - wt) + p2 * wt) >> 13);
                    } else {
                        p = 0;
                    }
                    cxt = u32(idx) + u32(wt >> 5);
                    sse_hi = (wt >> 5) != 0;
                    break;
                }
                default: break;
                }
            }
            // the MIX components of this level: a dot product over lanes each
            u32 mixers = __ballot_sync(kAll, now && type == C_MIX);
            while (mixers) {
                const int i = __ffs(int(mixers)) - 1;
                mixers &= mixers - 1;
                const i32 j = __shfl_sync(kAll, b, i), m = __shfl_sync(kAll, limit, i);
                const u32 base = __shfl_sync(kAll, cxt, i) * u32(m);
                const u64 tab = __shfl_sync(kAll, u64(reinterpret_cast<uintptr_t>(cm)), i);
                const i32 l = lane - j;
                i32 term = 0;
                if (l >= 0 && l < m && lane < n) {
                    const i32 w = i32(reinterpret_cast<const u32 *>(static_cast<uintptr_t>(tab))[base + u32(l)]) >> 8;
                    term = w * (lane < i ? p : pp);
                }
                const i32 sum = i32(__reduce_add_sync(kAll, u32(term)));
                if (lane == i) p = d_clamp2k(sum >> 8);
            }
        }
        return squash(__shfl_sync(kAll, p, n - 1));
    }

    __device__ void update(i32 y) {  // predictor.v:672-824
        const i32 t = y ? 32767 : 0;
        const i32 pj = __shfl_sync(kAll, p, ja), pk = __shfl_sync(kAll, p, jb);  // the finished p[]
        if (act) {
            switch (type) {
            case C_CM: {
                const u32 idx = cxt & (cm_len - 1);
                const u32 pn = u32(e0);  // cm[idx] as predict read it
                const i32 count = i32(pn & 0x3ff);
                const i32 err = t - i32(pn >> 17);
                const i32 upd = i32(u32(err) * u32(dt[count])) & -1024;  // wraps like V int
                cm[idx] = u32(i32(pn) + upd + (count < limit ? 1 : 0));
                break;
            }
            case C_ICM: {
                slot_set(hmap4 & 15u, cxt, nex[(cxt & 255u) * 2 + u32(y)]);
                const u32 v = cm[cxt];
                cm[cxt] = u32(i32(v) + ((t - i32(v >> 8)) >> 2));
                break;
            }
            case C_MATCH: {
                const i32 mask = i32(ht_len - 1);
                if (c != y) a = 0;
                const i32 idx = limit & mask;
                ht[idx] = u8((u32(ht[idx]) << 1) | u32(y));
                cxt++;
                if (cxt >= 8) {
                    cxt = 0;
                    limit = (limit + 1) & mask;
                    const i32 slot = i32(h) & i32(cm_len - 1);
                    if (a == 0) {
                        b = limit - i32(cm[slot]);
                        if ((b & mask) != 0) {
                            while (a < 255) {
                                const i32 i1 = (limit - a - 1) & mask;
                                const i32 i2 = (limit - a - b - 1) & mask;
                                if (ht[i1] != ht[i2]) break;
                                a++;
                            }
                        }
                    } else if (a < 255) {
                        a++;
                    }
                    cm[slot] = u32(limit);
                }
                break;
            }
            case C_MIX2: {
                const i32 err = ((t - squash(p)) * i32(q2)) >> 5;
                if (va && vb) {
                    const i32 w = e0 + ((err * (pj - pk) + 4096) >> 13);  // e0 = a16[cxt] as predict read it
                    a16[cxt] = u16(max(0, min(65535, w)));
                }
                break;
            }
            case C_ISSE: {
                const i32 err = t - squash(p);
                if (va) {
                    const i32 w0 = d_clamp512k(e0 + ((err * pj + 4096) >> 13));
                    const i32 w1 = d_clamp512k(e1 + ((err + 16) >> 5));
                    *reinterpret_cast<uint2 *>(cm + cxt * 2) = make_uint2(u32(w0), u32(w1));
                }
                slot_set(hmap4 & 15u, cxt, nex[(cxt & 255u) * 2 + u32(y)]);
                break;
            }
            case C_SSE: {
                const u32 idx = cxt & (cm_len - 1);
                // in range, cxt is idx or idx + 1 of predict: the cell is already in e0 / e1
                u32 v = sse_ok ? u32(sse_hi ? e1 : e0) : cm[idx];
                const i32 err = t - i32(v >> 17);
                const i32 count = i32(v) & 1023;
                if (count < limit) v = u32(i32(v) + ((err * (limit - count) + 4096) >> 13) + 1);
                cm[idx] = v;
                break;
            }
            default: break;
            }
        }
        // MIX: one weight per lane (predictor.v:763-775)
        u32 mixers = __ballot_sync(kAll, act && type == C_MIX);
        while (mixers) {
            const int i = __ffs(int(mixers)) - 1;
            mixers &= mixers - 1;
            const i32 err_own = ((t - squash(p)) * i32(q0)) >> 4;
            const i32 err = __shfl_sync(kAll, err_own, i);
            const i32 j = __shfl_sync(kAll, b, i), m = __shfl_sync(kAll, limit, i);
            const u32 base = __shfl_sync(kAll, cxt, i) * u32(m);
            const u64 tab = __shfl_sync(kAll, u64(reinterpret_cast<uintptr_t>(cm)), i);
            const i32 l = lane - j;
            if (l >= 0 && l < m && lane < n) {
                u32 *w = reinterpret_cast<u32 *>(static_cast<uintptr_t>(tab)) + base + u32(l);
                *w = u32(d_clamp512k(i32(*w) + ((err * p + 4096) >> 13)));
            }
        }
        __syncwarp();
        c8 = (c8 << 1) | u32(y);  // predictor.v:808-823
        if (c8 >= 256) {
            vm.run(c8 - 256);
            if (act && u32(lane) < h_len) h = vm.h[lane];
            hmap4 = 1, c8 = 1;
        } else if (c8 >= 16 && c8 < 32) {
            hmap4 = ((hmap4 & 0xf) << 5) | (u32(y) << 4) | 1;
        } else {
            hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + u32(y)) & 0xf);
        }
    }
};

__device__ __forceinline__ void load_const_tables(u8 *smem, const DevTables &T) {
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

// Coded bytes: every lane tracks the length, lane 0 stores.
struct SinkW {
    u8 *dst;
    u64 cap, len;
    int lane;
    __device__ __forceinline__ void put(u32 b) {
        if (lane == 0 && len < cap) dst[len] = u8(b);
        ++len;
    }
};
struct SourceW {
    const u8 *base;
    u64 pos, end;
    __device__ __forceinline__ i32 get() { return pos < end ? i32(base[pos++]) : -1; }
};

__device__ __forceinline__ void enc_bit(u32 &low, u32 &high, i32 y, u32 p16, SinkW &out) {  // encoder.v:48-89
    const u32 mid = coder_mid(low, high, p16);
    if (y) high = mid; else low = mid + 1;
    while ((high ^ low) < 0x1000000u) {
        out.put(high >> 24);
        low <<= 8;
        high = (high << 8) | 0xFFu;
        if (low == 0) low = 1;
    }
}
__device__ __forceinline__ i32 dec_bit(u32 &low, u32 &high, u32 &code, u32 p16, SourceW &in) {  // decoder.v:73-118
    const u32 mid = coder_mid(low, high, p16);
    i32 y;
    if (code <= mid) y = 1, high = mid; else y = 0, low = mid + 1;
    while ((high ^ low) < 0x1000000u) {
        low <<= 8;
        high = (high << 8) | 0xFFu;
        if (low == 0) low = 1;
        const i32 c = in.get();
        code = c < 0 ? (code << 8) : ((code << 8) | u32(c));
    }
    return y;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// k_encode_genwarp: Compressor.compress -> Encoder.compress -> Predictor.predict/update -> ZPAQL.run
// for every segment of every block of the wave (compressor.v:259-293, :375-378)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_encode_genwarp(EncodeArgs A) {
    extern __shared__ __align__(16) u8 smem[];
    load_const_tables(smem, A.tables);
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerCta + wic;
    if (warp >= A.n_blocks) return;
    GenW g;
    g.block_init(A.model, A.workspace + u64(warp) * A.model.ws_bytes, A.tables, smem,
                 smem + kConstBytes + size_t(wic) * kWarpBytes);
    const EncBlock blk = A.blocks[A.order[A.first_block + warp]];
    for (u32 s = 0; s < blk.n_seg; ++s) {
        const EncSeg seg = A.segs[blk.first_seg + s];
        SinkW out{A.arena + seg.pay_off, seg.pay_cap, 0, lane};
        g.segment_reset();
        u32 low = 1, high = 0xFFFFFFFFu;
        const u8 *src = A.in + seg.in_off;
        const u64 total = seg.in_len + ((seg.flags & 1u) ? 1u : 0u);
        for (u64 k = 0; k < total; ++k) {
            // the PP byte (0 = PASS) goes through the model first (compressor.v:271-274)
            const u32 ch = (seg.flags & 1u) ? (k == 0 ? 0u : src[k - 1]) : src[k];
            enc_bit(low, high, 0, 0, out);  // "not EOF" (encoder.v:108)
            for (int bit = 7; bit >= 0; --bit) {
                const i32 y = (ch >> bit) & 1;
                const i32 p = g.predict();
                enc_bit(low, high, y, u32(p * 2 + 1), out);
                g.update(y);
            }
        }
        enc_bit(low, high, 1, 0, out);  // EOF (encoder.v:101-105)
        out.put(high >> 24), out.put((high >> 16) & 255), out.put((high >> 8) & 255), out.put(high & 255);
        if (lane == 0) A.pay_len[blk.first_seg + s] = out.len;
    }
}

// ------------------------------------------------------------------------------------------
// k_decode_genwarp: find_filename / decompress / read_segment_end for every segment of a block
// (decompressor.v:350-635, decoder.v:29-196); PASS post-processing only.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_decode_genwarp(DecodeArgs A) {
    extern __shared__ __align__(16) u8 smem[];
    load_const_tables(smem, A.tables);
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerCta + wic;
    if (warp >= A.n_blocks) return;
    const int bi = int(A.order[A.first_block + warp]);
    GenW g;
    g.block_init(A.model, A.workspace + u64(warp) * A.model.ws_bytes, A.tables, smem,
                 smem + kConstBytes + size_t(wic) * kWarpBytes);
    const DecBlock blk = A.blocks[bi];
    SourceW in{A.arc, blk.arc_pos, A.arc_len};
    DecBlockOut res;
    res.end_pos = blk.arc_pos, res.out_len = 0, res.n_seg = 0, res.status = ZPAQGPU_OK;
    u8 *dst = A.out + blk.out_off;
    for (;;) {
        const i32 marker = in.get();  // decompressor.v:356-365
        if (marker < 0) { res.status = ZPAQGPU_E_FORMAT; break; }
        if (marker == 0xFF) break;
        DecSegRec rec;
        rec.block = u32(bi), rec.index = res.n_seg, rec.sha_off = ~0ull;
        rec.name_off = in.pos;
        i32 c;
        bool block_over = false;
        while ((c = in.get()) > 0)
            if (c == 0xFF) { block_over = true; break; }  // decompressor.v:380-384
        if (block_over) break;
        if (c < 0) { res.status = ZPAQGPU_E_FORMAT; break; }
        rec.comment_off = in.pos;
        while ((c = in.get()) > 0) {}
        if (c < 0 || in.get() < 0) { res.status = ZPAQGPU_E_FORMAT; break; }
        g.segment_reset();
        u32 low = 1, high = 0xFFFFFFFFu, code = 0;
        for (int k = 0; k < 4; ++k) {  // decoder.v:37-46
            const i32 b = in.get();
            code = b < 0 ? (code << 8) : ((code << 8) | u32(b));
        }
        rec.out_off = blk.out_off + res.out_len;
        u64 produced = 0;
        int pp_state = 0;
        for (;;) {
            if (dec_bit(low, high, code, 0, in)) break;  // EOF flag (decoder.v:128-131)
            u32 ch = 1;
            while (ch < 256) {
                const i32 p = g.predict();
                const i32 y = dec_bit(low, high, code, u32(p * 2 + 1), in);
                g.update(y);
                ch = (ch << 1) | u32(y);
            }
            ch -= 256;
            if (pp_state == 0) {  // PostProcessor.write state 0 (decompressor.v:58-70)
                pp_state = (ch + 1 > 2) ? 1 : i32(ch) + 1;
                if (pp_state == 2) { res.status = ZPAQGPU_E_UNSUPPORTED; break; }
            } else {
                const u64 at = res.out_len + produced;
                if (lane == 0 && at < blk.out_cap) dst[at] = u8(ch);
                ++produced;
            }
        }
        if (res.status != ZPAQGPU_OK) break;
        // Decoder.skip (decoder.v:151-196) then read_segment_end (decompressor.v:608-631)
        u32 curr = code;
        i32 mk = 0;
        bool eof = false;
        if (curr == 0) {
            const i32 b = in.get();
            if (b < 0) eof = true; else curr = u32(b);
        }
        while (!eof && curr != 0) {
            const i32 b = in.get();
            if (b < 0) eof = true; else curr = (curr << 8) | u32(b);
        }
        while (!eof) {
            mk = in.get();
            if (mk < 0) eof = true;
            if (mk != 0) break;
        }
        if (!eof && mk == 253) {
            rec.sha_off = in.pos;
            in.pos = min(in.pos + 20, in.end);
        }
        rec.out_len = produced;
        res.out_len += produced;
        if (lane == 0) {
            const u32 slot = atomicAdd(A.seg_count, 1u);
            if (slot < A.seg_cap) A.seg_recs[slot] = rec;
        }
        res.n_seg++;
    }
    res.end_pos = in.pos;
    if (lane == 0) A.results[bi] = res;
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
bool genwarp_supports(const Model &m) { return m.n >= 1 && m.n <= 32; }

static size_t genwarp_smem() { return kConstBytes + size_t(kWarpsPerCta) * kWarpBytes; }

bool launch_encode_genwarp(const EncodeArgs &A, cudaStream_t s) {
    if (cudaFuncSetAttribute(k_encode_genwarp, cudaFuncAttributeMaxDynamicSharedMemorySize, int(genwarp_smem())) !=
        cudaSuccess)
        return false;
    k_encode_genwarp<<<(A.n_blocks + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, genwarp_smem(), s>>>(A);
    return true;
}
bool launch_decode_genwarp(const DecodeArgs &A, cudaStream_t s) {
    if (cudaFuncSetAttribute(k_decode_genwarp, cudaFuncAttributeMaxDynamicSharedMemorySize, int(genwarp_smem())) !=
        cudaSuccess)
        return false;
    k_decode_genwarp<<<(A.n_blocks + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, genwarp_smem(), s>>>(A);
    return true;
}

}  // namespace zg

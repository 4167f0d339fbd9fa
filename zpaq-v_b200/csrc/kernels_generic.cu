// kernels_generic.cu -- the all-components codec kernel: any header the reference decoder accepts
// (decompressor.v:278-342): CONS/CM/ICM/MATCH/AVG/MIX2/MIX/ISSE/SSE in any wiring plus an
// arbitrary HCOMP program run by a ZPAQL interpreter.
//
// Mapping: one ZPAQ block per warp, executed by lane 0 (a block is a strictly serial bit stream;
// the other lanes exit).  All model state lives in the block's HBM workspace.  This is the
// catch-all; the predefined -m1..-m5 shapes go to the specialised kernel in kernels_chain.cu.
#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"

namespace zg {
namespace {

// ------------------------------------------------------------------------------------------
// ZPAQL interpreter (zpaql.v:167-954), registers in thread registers, M/H/R in the workspace.
// ------------------------------------------------------------------------------------------
struct Vm {
    u32 a, b, c, d;
    i32 f, pc;
    u8 *m;
    u32 *h;
    u32 *r;
    u32 m_mask, h_mask;  // len-1; arrays of length 0 are flagged by has_m / has_h
    bool has_m, has_h;
    const u8 *hdr;
    i32 hbegin, hend, hdr_len;

    __device__ u32 mget(u32 i) const { return has_m ? m[i & m_mask] : 0u; }
    __device__ void mset(u32 i, u32 v) {
        if (has_m) m[i & m_mask] = u8(v);
    }
    __device__ u32 hget(u32 i) const { return has_h ? h[i & h_mask] : 0u; }
    __device__ void hset(u32 i, u32 v) {
        if (has_h) h[i & h_mask] = v;
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

    // One instruction; false stops the run (zpaql.v:215-954).
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
            switch ((op >> 3) & 7) {
            case 0: a = v; break;
            case 1: b = v; break;
            case 2: c = v; break;
            case 3: d = v; break;
            case 4: mset(b, v); break;
            case 5: mset(c, v); break;
            default: hset(d, v); break;
            }
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
        if (op < 32 && (op & 7) < 5) {  // register group ops: swap/++/--/~/=0 on a,b,c,d
            u32 *reg = (op < 8) ? &a : (op < 16) ? &b : (op < 24) ? &c : &d;
            switch (op & 7) {
            case 0:
                if (op) {  // op 0 is NOP; 8/16/24 swap with a
                    const u32 t = a;
                    a = *reg, *reg = t;
                }
                break;
            case 1: ++*reg; break;
            case 2: --*reg; break;
            case 3: *reg = ~*reg; break;
            default: *reg = 0; break;
            }
            return true;
        }
        switch (op) {
        case 7: a = r[operand & 255]; break;
        case 15: b = r[operand & 255]; break;
        case 23: c = r[operand & 255]; break;
        case 31: d = r[operand & 255]; break;
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
        case 55: r[operand & 255] = a; break;
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
// Predictor over the workspace (predictor.v:495-824)
// ------------------------------------------------------------------------------------------
struct Gen {
    const ModelDev *M;
    DevTables T;
    u8 *ws;
    CompRt *rt;
    Vm vm;
    u32 c8, hmap4;
    i32 n;

    __device__ i32 squash(i32 d) const { return T.squash[d_squash_idx(d)]; }
    __device__ i32 stretch(i32 p) const { return T.stretch[d_stretch_idx(p)]; }
    __device__ i32 nex(i32 s, i32 y) const { return T.nex[(s & 255) * 2 + y]; }

    __device__ void block_init() {
        n = M->n;
        rt = reinterpret_cast<CompRt *>(ws + M->rt_off);
        for (i32 i = 0; i < n; ++i) {
            const CompDesc &cd = M->comps[i];
            CompRt z;
            z.a = cd.a, z.b = cd.b, z.c = cd.c, z.limit = cd.limit, z.cxt = 0, z.p = 0, z.h = 0, z.pad = 0;
            rt[i] = z;
        }
        vm.a = vm.b = vm.c = vm.d = 0, vm.f = 0, vm.pc = M->hbegin;
        vm.m = ws + M->m_off, vm.h = reinterpret_cast<u32 *>(ws + M->h_off);
        vm.r = reinterpret_cast<u32 *>(ws + M->r_off);
        vm.has_m = M->m_len != 0, vm.has_h = M->h_len != 0;
        vm.m_mask = M->m_len - 1, vm.h_mask = M->h_len - 1;
        vm.hdr = M->header, vm.hbegin = M->hbegin, vm.hend = M->hend, vm.hdr_len = M->header_len;
        c8 = 1, hmap4 = 1;
    }
    __device__ void segment_reset() {  // predictor.v:827-833
        c8 = 1, hmap4 = 1;
        for (i32 i = 0; i < n; ++i) rt[i].h = 0;
    }

    // predictor.v:495-532
    __device__ i32 find_slot(u8 *ht, u32 ht_len, i32 sizebits, u32 cxt) const {
        const u32 chk = (cxt >> sizebits) & 255u;
        const u32 h0 = (cxt * 16u) & (ht_len - 16u), h1 = h0 ^ 16u, h2 = h0 ^ 32u;
        if (ht[h0] == chk) return i32(h0);
        if (ht[h1] == chk) return i32(h1);
        if (ht[h2] == chk) return i32(h2);
        const u32 q0 = ht[h0 + 1], q1 = ht[h1 + 1], q2 = ht[h2 + 1];
        const u32 victim = (q0 <= q1 && q0 <= q2) ? h0 : (q1 < q2 ? h1 : h2);
        uint4 *slot = reinterpret_cast<uint4 *>(ht + victim);
        *slot = make_uint4(chk, 0u, 0u, 0u);
        return i32(victim);
    }

    __device__ i32 predict() {  // predictor.v:536-668
        if (n == 0) return 16384;
        for (i32 i = 0; i < n; ++i) {
            const CompDesc &cd = M->comps[i];
            CompRt &cr = rt[i];
            u32 *cm = reinterpret_cast<u32 *>(ws + cd.cm_off);
            u8 *ht = ws + cd.ht_off;
            i32 p = 0;
            switch (cd.type) {
            case C_CONS: p = (cr.a - 128) * 16; break;
            case C_CM: {
                cr.cxt = cr.h ^ hmap4;
                p = stretch(i32(cm[i32(cr.cxt) & i32(cd.cm_len - 1)] >> 17));
                break;
            }
            case C_ICM: {
                if (c8 == 1 || (c8 & 0xf0) == 16) cr.c = find_slot(ht, cd.ht_len, cr.a + 2, cr.h + 16u * c8);
                cr.cxt = ht[cr.c + i32(hmap4 & 15)];
                p = stretch(i32(cm[cr.cxt] >> 8));
                break;
            }
            case C_MATCH: {
                if (cr.a == 0) {
                    p = 0;
                } else {
                    const i32 idx = (cr.limit - cr.b) & i32(cd.ht_len - 1);
                    cr.c = i32((u32(ht[idx]) >> (7 - i32(cr.cxt))) & 1u);
                    p = stretch((T.dt2k[cr.a & 255] * (cr.c * -2 + 1)) & 32767);
                }
                break;
            }
            case C_AVG: {
                const i32 j = cr.a, k = cr.b, wt = cr.c;
                p = (j < n && k < n) ? ((rt[j].p * wt + rt[k].p * (256 - wt)) >> 8) : 0;
                break;
            }
            case C_MIX2: {
                const i32 j = i32(cd.p[0]), k = i32(cd.p[1]);
                cr.cxt = (cr.h + (c8 & cd.p[3])) & u32(cr.c - 1);
                const i32 w = reinterpret_cast<const u16 *>(ws + cd.a16_off)[cr.cxt];
                p = (j < n && k < n) ? d_clamp2k((w * rt[j].p + (65536 - w) * rt[k].p) >> 16) : 0;
                break;
            }
            case C_MIX: {
                const i32 j = cr.b, m = cr.limit, mask = i32(cd.p[1]);
                cr.cxt = u32((i32(cr.h) + (i32(c8) & mask)) & (cr.c - 1));
                const i32 base = i32(cr.cxt) * m;
                i32 sum = 0;
                for (i32 l = 0; l < m && (j + l) < n; ++l) sum += (i32(cm[base + l]) >> 8) * rt[j + l].p;
                p = d_clamp2k(sum >> 8);
                break;
            }
            case C_ISSE: {
                if (c8 == 1 || (c8 & 0xf0) == 16) cr.c = find_slot(ht, cd.ht_len, cr.a + 2, cr.h + 16u * c8);
                cr.cxt = ht[cr.c + i32(hmap4 & 15)];
                const i32 w0 = i32(cm[cr.cxt * 2]), w1 = i32(cm[cr.cxt * 2 + 1]);
                const i32 j = cr.b;
                p = (j < n) ? d_clamp2k((w0 * rt[j].p + w1 * 64) >> 16) : d_clamp2k(w1 >> 10);
                break;
            }
            case C_SSE: {
                const i32 j = cr.b;
                cr.cxt = (cr.h + c8) * 32u;
                i32 pq = (j < n) ? rt[j].p + 992 : 992;
                pq = max(0, min(1983, pq));
                const i32 wt = pq & 63;
                pq >>= 6;
                const i32 idx = i32(cr.cxt) + pq;
                if (idx >= 0 && idx + 1 < i32(cd.cm_len)) {
                    const i32 p1 = i32(cm[idx] >> 10), p2 = i32(cm[idx + 1] >> 10);
                    p = stretch((p1 * (64 - wt) + p2 * wt) >> 13);
                } else {
                    p = 0;
                }
                cr.cxt = u32(idx) + u32(wt >> 5);
                break;
            }
            default: p = 0; break;
            }
            cr.p = p;
        }
        return squash(rt[n - 1].p);
    }

    __device__ void update(i32 y) {  // predictor.v:672-824
        const i32 t = y ? 32767 : 0;
        for (i32 i = 0; i < n; ++i) {
            const CompDesc &cd = M->comps[i];
            CompRt &cr = rt[i];
            u32 *cm = reinterpret_cast<u32 *>(ws + cd.cm_off);
            u8 *ht = ws + cd.ht_off;
            switch (cd.type) {
            case C_CM: {
                const i32 idx = i32(cr.cxt) & i32(cd.cm_len - 1);
                const u32 pn = cm[idx];
                const i32 count = i32(pn & 0x3ff);
                const i32 err = t - i32(pn >> 17);
                const i32 upd = i32(u32(err) * u32(T.dt[count])) & -1024;  // wraps like V int
                cm[idx] = u32(i32(pn) + upd + (count < cr.limit ? 1 : 0));
                break;
            }
            case C_ICM: {
                const i32 at = cr.c + i32(hmap4 & 15);
                ht[at] = u8(nex(ht[at], y));
                const u32 v = cm[cr.cxt];
                cm[cr.cxt] = u32(i32(v) + ((t - i32(v >> 8)) >> 2));
                break;
            }
            case C_MATCH: {
                const i32 mask = i32(cd.ht_len - 1);
                if (cr.c != y) cr.a = 0;
                const i32 idx = cr.limit & mask;
                ht[idx] = u8((u32(ht[idx]) << 1) | u32(y));
                cr.cxt++;
                if (cr.cxt >= 8) {
                    cr.cxt = 0;
                    cr.limit = (cr.limit + 1) & mask;
                    const i32 slot = i32(cr.h) & i32(cd.cm_len - 1);
                    if (cr.a == 0) {
                        cr.b = cr.limit - i32(cm[slot]);
                        if ((cr.b & mask) != 0) {
                            while (cr.a < 255) {
                                const i32 i1 = (cr.limit - cr.a - 1) & mask;
                                const i32 i2 = (cr.limit - cr.a - cr.b - 1) & mask;
                                if (ht[i1] != ht[i2]) break;
                                cr.a++;
                            }
                        }
                    } else if (cr.a < 255) {
                        cr.a++;
                    }
                    cm[slot] = u32(cr.limit);
                }
                break;
            }
            case C_MIX2: {
                const i32 j = i32(cd.p[0]), k = i32(cd.p[1]), rate = i32(cd.p[2]);
                const i32 err = ((t - squash(cr.p)) * rate) >> 5;
                if (j < n && k < n) {
                    u16 *a16 = reinterpret_cast<u16 *>(ws + cd.a16_off);
                    i32 w = a16[cr.cxt];
                    w += (err * (rt[j].p - rt[k].p) + 4096) >> 13;
                    a16[cr.cxt] = u16(max(0, min(65535, w)));
                }
                break;
            }
            case C_MIX: {
                const i32 j = cr.b, m = cr.limit, rate = i32(cd.p[0]);
                const i32 err = ((t - squash(cr.p)) * rate) >> 4;
                const i32 base = i32(cr.cxt) * m;
                for (i32 l = 0; l < m && (j + l) < n; ++l)
                    cm[base + l] = u32(d_clamp512k(i32(cm[base + l]) + ((err * rt[j + l].p + 4096) >> 13)));
                break;
            }
            case C_ISSE: {
                const i32 j = cr.b;
                const i32 err = t - squash(cr.p);
                if (j < n) {
                    const i32 w0 = d_clamp512k(i32(cm[cr.cxt * 2]) + ((err * rt[j].p + 4096) >> 13));
                    const i32 w1 = d_clamp512k(i32(cm[cr.cxt * 2 + 1]) + ((err + 16) >> 5));
                    cm[cr.cxt * 2] = u32(w0), cm[cr.cxt * 2 + 1] = u32(w1);
                }
                ht[cr.c + i32(hmap4 & 15)] = u8(nex(i32(cr.cxt), y));
                break;
            }
            case C_SSE: {
                const i32 idx = i32(cr.cxt) & i32(cd.cm_len - 1);
                u32 v = cm[idx];
                const i32 err = t - i32(v >> 17);
                const i32 count = i32(v) & 1023;
                if (count < cr.limit) v = u32(i32(v) + ((err * (cr.limit - count) + 4096) >> 13) + 1);
                cm[idx] = v;
                break;
            }
            default: break;
            }
        }
        c8 = (c8 << 1) | u32(y);  // predictor.v:808-823
        if (c8 >= 256) {
            vm.run(c8 - 256);
            for (i32 i = 0; i < n && u32(i) < M->h_len; ++i) rt[i].h = vm.h[i];
            hmap4 = 1, c8 = 1;
        } else if (c8 >= 16 && c8 < 32) {
            hmap4 = ((hmap4 & 0xf) << 5) | (u32(y) << 4) | 1;
        } else {
            hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + u32(y)) & 0xf);
        }
    }
};

// Byte sink into the payload slot: counts everything, stores while capacity lasts.
struct Sink {
    u8 *dst;
    u64 cap, len;
    __device__ void put(u32 b) {
        if (len < cap) dst[len] = u8(b);
        ++len;
    }
};
// Byte source over the archive.
struct Source {
    const u8 *base;
    u64 pos, end;
    __device__ i32 get() { return pos < end ? i32(base[pos++]) : -1; }
};

__device__ void enc_bit(u32 &low, u32 &high, i32 y, u32 p16, Sink &out) {  // encoder.v:48-89
    const u32 mid = coder_mid(low, high, p16);
    if (y) high = mid; else low = mid + 1;
    while ((high ^ low) < 0x1000000u) {
        out.put(high >> 24);
        low <<= 8;
        high = (high << 8) | 0xFFu;
        if (low == 0) low = 1;
    }
}
__device__ i32 dec_bit(u32 &low, u32 &high, u32 &code, u32 p16, Source &in) {  // decoder.v:73-118
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
// k_encode_generic: Compressor.compress -> Encoder.compress -> Predictor.predict/update ->
// ZPAQL.run for every segment of every block of the wave (compressor.v:259-293, :375-378).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_encode_generic(EncodeArgs A) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= A.n_blocks || (threadIdx.x & 31) != 0) return;
    Gen g;
    g.M = &A.model, g.T = A.tables;
    g.ws = A.workspace + u64(warp) * A.model.ws_bytes;
    g.block_init();
    const EncBlock blk = A.blocks[A.order[A.first_block + warp]];
    for (u32 s = 0; s < blk.n_seg; ++s) {
        const EncSeg seg = A.segs[blk.first_seg + s];
        Sink out{A.arena + seg.pay_off, seg.pay_cap, 0};
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
        A.pay_len[blk.first_seg + s] = out.len;
    }
}

// ------------------------------------------------------------------------------------------
// k_decode_generic: find_filename / decompress / read_segment_end for every segment of a block
// (decompressor.v:350-635, decoder.v:29-196); PASS post-processing only.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_decode_generic(DecodeArgs A) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= A.n_blocks || (threadIdx.x & 31) != 0) return;
    const int bi = int(A.order[A.first_block + warp]);
    Gen g;
    g.M = &A.model, g.T = A.tables;
    g.ws = A.workspace + u64(warp) * A.model.ws_bytes;
    g.block_init();
    const DecBlock blk = A.blocks[bi];
    Source in{A.arc, blk.arc_pos, A.arc_len};
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
                if (at < blk.out_cap) dst[at] = u8(ch);
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
        const u32 slot = atomicAdd(A.seg_count, 1u);
        if (slot < A.seg_cap) A.seg_recs[slot] = rec;
        res.n_seg++;
    }
    res.end_pos = in.pos;
    A.results[bi] = res;
}

}  // namespace zg

// kernels_aux.cu -- the kernels either side of the codec kernel: table initialisation, SHA-1 of
// segment plaintext, block assembly, the block locator scan and store-mode decoding.
#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"
#include "sha1_lane.h"

#include <cstdlib>
#include <cstring>

namespace zg {

// ------------------------------------------------------------------------------------------
// k_fill_workspace: initial images of the adaptive tables (predictor.v:352-354, :366-368,
// :396, :425-427, :442-446, :459-462) replicated into every workspace slot.  Hash tables and
// ZPAQL memory start as zero and are cleared by a memset of the whole workspace.
// ------------------------------------------------------------------------------------------
__global__ void k_fill_workspace(FillArgs A) {
    const int slot = blockIdx.y;
    const FillRegion r = A.regions[blockIdx.z];
    u32 *dst = reinterpret_cast<u32 *>(A.workspace + u64(slot) * A.ws_bytes + r.off);
    const u32 *img = A.image + r.img_off;
    for (u64 k = u64(blockIdx.x) * blockDim.x + threadIdx.x; k < r.n_words; k += u64(gridDim.x) * blockDim.x)
        dst[k] = img[r.period == 1 ? 0 : k % r.period];
}

void launch_fill(const FillArgs &A, cudaStream_t s) {
    if (A.n_regions == 0 || A.n_slots == 0) return;
    // regions are small except CM/MIX/SSE tables; a handful of CTAs per (slot, region) suffices
    // and large tables grid-stride.  gridDim.y is limited to 65535 slots per launch.
    for (int first = 0; first < A.n_slots; first += 65535) {
        FillArgs B = A;
        B.workspace = A.workspace + u64(first) * A.ws_bytes;
        const int n = A.n_slots - first < 65535 ? A.n_slots - first : 65535;
        dim3 grid(8, unsigned(n), unsigned(A.n_regions));
        k_fill_workspace<<<grid, 256, 0, s>>>(B);
    }
}

// ------------------------------------------------------------------------------------------
// k_sha1_segments: SHA-1 of each byte range (sha1.v:42-146), one range per lane, one warp per CTA.
//
// A hash is one serial chain (80 rounds of two dependent operations per 64 bytes), so a range belongs to one
// lane and the kernel's time is that of its longest range; what the warp shares is the memory side.  The
// ranges start at any byte (jidac fragments) and lie megabytes apart, so a lane loading its own range would
// issue 32 separate sector requests per instruction and, unaligned, one per BYTE.  Instead the warp brings
// 272 bytes (four blocks and the 16 bytes an unaligned block can spill into) of every lane's range into
// shared memory with 16-byte cp.async copies -- lane t of the warp copies chunk t of a row, so a row is one
// coalesced request -- two rounds deep, so that the copies of round r + 1 fly while round r is hashed.  The
// lane then takes its 16 big-endian words from its row with PRMT at whatever alignment the range has, and
// makes the padding blocks from the same window (sha1_lane.h; the identical code runs under a host emulation
// of the warp in tests/test_sha1_lane.py).  Nothing outside a range is ever read: the first and last chunk
// of an unaligned range are copied byte by byte.
// ------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ void cp_async16(void *smem_dst, u64 gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(u32(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
}  // namespace

__global__ void __launch_bounds__(32) k_sha1_segments(const u8 *base, const ShaJob *jobs, int n_jobs, u8 *digests) {
    using namespace sha1lane;
    __shared__ __align__(16) u8 buf[2][32 * kRowBytes];
    __shared__ u64 desc[32][3];   // per row: window origin, first byte, one past the last
    const int lane = threadIdx.x;
    const int j = blockIdx.x * 32 + lane;
    const bool has = j < n_jobs;
    Lane L;
    lane_begin(L, reinterpret_cast<u64>(base) + (has ? jobs[j].off : 0ull), has ? jobs[j].len : 0ull, has);
    desc[lane][0] = L.a, desc[lane][1] = L.s, desc[lane][2] = L.end;
    const u32 n_rounds = __reduce_max_sync(0xFFFFFFFFu, lane_rounds(L));
    __syncwarp();
    auto issue = [&](u32 r) {
        u8 *stage = buf[r & 1];
#pragma unroll 1
        for (u32 i = 0; i < kChunksPerRow; ++i) {
            const u32 q = u32(lane) + 32u * i;
            const u32 row = q / kChunksPerRow, c = q - row * kChunksPerRow;
            u64 src;
            u32 lo = 0, hi = 0;
            const ChunkKind kind = chunk_plan(desc[row][0], desc[row][1], desc[row][2], r, c, src, lo, hi);
            u8 *dst = stage + row * kRowBytes + 16u * c;
            if (kind == kWhole) {
                cp_async16(dst, src);
            } else if (kind == kPart) {
                for (u32 t = lo; t < hi; ++t) dst[t] = *reinterpret_cast<const u8 *>(src + t);
            }
        }
    };
    if (n_rounds) issue(0);
    cp_async_commit();
    for (u32 r = 0; r < n_rounds; ++r) {
        if (r + 1 < n_rounds) issue(r + 1);
        cp_async_commit();            // (an empty group on the last round keeps the count uniform)
        cp_async_wait_all_but_one();  // round r has landed; round r + 1 may still be in flight
        __syncwarp();
        const u32 *row = reinterpret_cast<const u32 *>(buf[r & 1] + lane * kRowBytes);
#pragma unroll 1
        for (u32 k = 0; k < kBlocksPerRound; ++k) {
            if (r * kBlocksPerRound + k < L.total) {
                u32 w[16];
                block_words(L, row, r, k, w);
                compress(L.st, w);
            }
        }
        __syncwarp();                 // the stage is free for round r + 2
    }
    if (has) digest_bytes(L, digests + u64(j) * 20);
}

// ------------------------------------------------------------------------------------------
// k_sha1_direct: the kernel k_sha1_segments replaced -- every lane loads its own range (16-byte loads when
// the range is aligned, bytes otherwise).  Kept as the A/B arm (ZPAQGPU_SHA1=direct).
// ------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ u32 rol(u32 x, int n) { return __funnelshift_l(x, x, n); }

__device__ void sha1_compress(u32 st[5], const u32 blockw[16]) {
    u32 w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = blockw[i];
    u32 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4];
#pragma unroll
    for (int i = 0; i < 80; ++i) {
        u32 wi;
        if (i < 16) {
            wi = w[i];
        } else {
            wi = rol(w[(i - 3) & 15] ^ w[(i - 8) & 15] ^ w[(i - 14) & 15] ^ w[i & 15], 1);
            w[i & 15] = wi;
        }
        u32 f, k;
        if (i < 20) f = (b & c) | (~b & d), k = 0x5A827999u;
        else if (i < 40) f = b ^ c ^ d, k = 0x6ED9EBA1u;
        else if (i < 60) f = (b & c) | (b & d) | (c & d), k = 0x8F1BBCDCu;
        else f = b ^ c ^ d, k = 0xCA62C1D6u;
        const u32 t = rol(a, 5) + f + e + k + wi;
        e = d, d = c, c = rol(b, 30), b = a, a = t;
    }
    st[0] += a, st[1] += b, st[2] += c, st[3] += d, st[4] += e;
}
}  // namespace

__global__ void k_sha1_direct(const u8 *base, const ShaJob *jobs, int n_jobs, u8 *digests) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const u8 *p = base + jobs[j].off;
    const u64 len = jobs[j].len;
    u32 st[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    u32 w[16];
    const u64 full = len / 64;
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    for (u64 blk = 0; blk < full; ++blk) {
        const u8 *q = p + blk * 64;
        if (aligned) {
            const uint4 *v = reinterpret_cast<const uint4 *>(q);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 x = v[i];
                w[i * 4 + 0] = __byte_perm(x.x, 0, 0x0123), w[i * 4 + 1] = __byte_perm(x.y, 0, 0x0123);
                w[i * 4 + 2] = __byte_perm(x.z, 0, 0x0123), w[i * 4 + 3] = __byte_perm(x.w, 0, 0x0123);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                w[i] = (u32(q[i * 4]) << 24) | (u32(q[i * 4 + 1]) << 16) | (u32(q[i * 4 + 2]) << 8) | u32(q[i * 4 + 3]);
        }
        sha1_compress(st, w);
    }
    // padding (sha1.v:112-134)
    u8 tail[128];
    const u32 rem = u32(len - full * 64);
    for (u32 i = 0; i < rem; ++i) tail[i] = p[full * 64 + i];
    u32 n = rem;
    tail[n++] = 0x80;
    const u32 padded = (n > 56) ? 128 : 64;
    while (n < padded - 8) tail[n++] = 0;
    const u64 bits = len * 8;
    for (int i = 7; i >= 0; --i) tail[n++] = u8(bits >> (i * 8));
    for (u32 o = 0; o < padded; o += 64) {
        for (int i = 0; i < 16; ++i)
            w[i] = (u32(tail[o + i * 4]) << 24) | (u32(tail[o + i * 4 + 1]) << 16) |
                   (u32(tail[o + i * 4 + 2]) << 8) | u32(tail[o + i * 4 + 3]);
        sha1_compress(st, w);
    }
    u8 *out = digests + u64(j) * 20;
    for (int i = 0; i < 5; ++i) {
        out[i * 4] = u8(st[i] >> 24), out[i * 4 + 1] = u8(st[i] >> 16);
        out[i * 4 + 2] = u8(st[i] >> 8), out[i * 4 + 3] = u8(st[i]);
    }
}


void launch_sha1(const u8 *base, const ShaJob *jobs, int n_jobs, u8 *digests, cudaStream_t s) {
    if (n_jobs <= 0) return;
    // one warp per CTA spreads the serial hashes over as many SMs as possible
    static const bool direct = [] {
        const char *v = std::getenv("ZPAQGPU_SHA1");
        return v && std::strcmp(v, "direct") == 0;
    }();
    if (direct) k_sha1_direct<<<(n_jobs + 31) / 32, 32, 0, s>>>(base, jobs, n_jobs, digests);
    else k_sha1_segments<<<(n_jobs + 31) / 32, 32, 0, s>>>(base, jobs, n_jobs, digests);
}

// ------------------------------------------------------------------------------------------
// Block assembly: sizes -> exclusive scan -> copy.
// ------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ u64 store_payload_bytes(u64 in_len, u32 flags) {
    // compressor.v:297-354: the PP byte and the data go out in chunks closed at 65536 bytes,
    // each behind a 4-byte big-endian length
    const u64 n = in_len + (flags & 1u);
    if (n == 0) return 0;
    return n + 4 * ((n + 65535) / 65536);
}
}  // namespace

__global__ void k_seg_sizes(PackArgs A, int n_segs) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const PackSeg g = A.segs[s];
    const u64 pay = g.store ? store_payload_bytes(g.in_len, g.flags) : A.pay_len[s];
    A.seg_size[s] = g.pre_len + pay + 25 + (g.last ? 1 : 0);  // 00000000 FD sha1[20] [FF]
}

// single CTA: block sizes and their exclusive prefix sum
__global__ void k_scan_blocks(PackArgs A) {
    __shared__ u64 partial[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (A.n_blocks + nt - 1) / nt;
    const int lo = min(A.n_blocks, t * per), hi = min(A.n_blocks, lo + per);
    u64 sum = 0;
    for (int b = lo; b < hi; ++b) {
        const EncBlock blk = A.blocks[b];
        for (u32 s = 0; s < blk.n_seg; ++s) sum += A.seg_size[blk.first_seg + s];
    }
    partial[t] = sum;
    __syncthreads();
    if (t == 0) {
        u64 run = 0;
        for (int i = 0; i < nt; ++i) {
            const u64 v = partial[i];
            partial[i] = run;
            run += v;
        }
        A.out_off[A.n_blocks] = run;
    }
    __syncthreads();
    u64 at = partial[t];
    for (int b = lo; b < hi; ++b) {
        A.out_off[b] = at;
        const EncBlock blk = A.blocks[b];
        for (u32 s = 0; s < blk.n_seg; ++s) at += A.seg_size[blk.first_seg + s];
    }
}

__global__ void k_pack_blocks(PackArgs A) {
    const int b = blockIdx.x;
    const EncBlock blk = A.blocks[b];
    u64 at = A.out_off[b];
    if (A.out_off[b + 1] > A.out_cap) return;  // does not fit: the host reports NOSPACE
    const int t = threadIdx.x, nt = blockDim.x;
    for (u32 si = 0; si < blk.n_seg; ++si) {
        const u32 s = blk.first_seg + si;
        const PackSeg g = A.segs[s];
        u8 *o = A.out + at;
        const u8 *pre = A.pre + g.pre_off;
        for (u32 i = t; i < g.pre_len; i += nt) o[i] = pre[i];
        o += g.pre_len;
        u64 pay;
        if (g.store) {
            const u64 n = g.in_len + (g.flags & 1u);
            pay = store_payload_bytes(g.in_len, g.flags);
            const u8 *src = A.in + g.in_off;
            const u64 pp = g.flags & 1u;
            for (u64 i = t; i < n; i += nt) {
                const u64 chunk = i >> 16;
                o[i + 4 * (chunk + 1)] = (i < pp) ? u8(0) : src[i - pp];
            }
            const u64 chunks = (n + 65535) / 65536;
            for (u64 c = t; c < chunks; c += nt) {
                const u64 len = (c + 1 < chunks) ? 65536 : n - c * 65536;
                u8 *h = o + c * 65540;
                h[0] = u8(len >> 24), h[1] = u8(len >> 16), h[2] = u8(len >> 8), h[3] = u8(len);
            }
        } else {
            pay = A.pay_len[s];
            const u8 *src = A.arena + g.pay_off;
            const u64 have = pay < g.pay_cap ? pay : g.pay_cap;  // an overflowed slot is retried by the host
            for (u64 i = t; i < have; i += nt) o[i] = src[i];
        }
        o += pay;
        if (t < 4) o[t] = 0;
        if (t == 4) o[4] = 253;
        if (t >= 5 && t < 25) o[t] = A.digests[u64(s) * 20 + (t - 5)];
        if (g.last && t == 25) o[25] = 0xFF;
        at += A.seg_size[s];
    }
}

void launch_pack(const PackArgs &A, int n_segs, cudaStream_t s) {
    if (A.n_blocks <= 0) return;
    k_seg_sizes<<<(n_segs + 255) / 256, 256, 0, s>>>(A, n_segs);
    k_scan_blocks<<<1, 1024, 0, s>>>(A);
    k_pack_blocks<<<A.n_blocks, 256, 0, s>>>(A);
}

// ------------------------------------------------------------------------------------------
// k_find_blocks: the four rolling hashes of find_block (decompressor.v:227-254).  After 16 bytes
// the start values have been multiplied away (12, 20, 28, 44 are multiples of 4, so m^16 = 0 mod
// 2^32); each thread therefore recomputes the hashes of position i from the start values and the
// last min(16, i+1) bytes, which is exact for every i.
// ------------------------------------------------------------------------------------------
__global__ void k_find_blocks(const u8 *arc, u64 len, u64 *starts, u32 cap, u32 *count) {
    const u64 i = u64(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= len) return;
    u32 h1 = 0x3D49B113u, h2 = 0x29EB7F93u, h3 = 0x2614BE13u, h4 = 0x3828EB13u;
    const u64 first = i >= 15 ? i - 15 : 0;
    for (u64 k = first; k <= i; ++k) {
        const u32 c = arc[k];
        h1 = h1 * 12u + c, h2 = h2 * 20u + c, h3 = h3 * 28u + c, h4 = h4 * 44u + c;
    }
    if (h1 == 0xB16B88F1u && h2 == 0xFF5376F1u && h3 == 0x72AC5BF1u && h4 == 0x2F909AF1u) {
        const u32 at = atomicAdd(count, 1u);
        if (at < cap) starts[at] = i + 1;
    }
}

void launch_find_blocks(const u8 *arc, u64 len, u64 *starts, u32 cap, u32 *count, cudaStream_t s) {
    if (len == 0) return;
    k_find_blocks<<<unsigned((len + 255) / 256), 256, 0, s>>>(arc, len, starts, cap, count);
}

// ------------------------------------------------------------------------------------------
// k_decode_store: blocks without components (decompress_store, decompressor.v:518-587), one warp
// per block: lane-uniform walk over the chunk lengths, whole-warp copies.
// ------------------------------------------------------------------------------------------
__global__ void k_decode_store(DecodeArgs A) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= A.n_blocks) return;
    const int bi = int(A.order[A.first_block + warp]);
    const DecBlock blk = A.blocks[bi];
    const u8 *arc = A.arc;
    u64 pos = blk.arc_pos;
    DecBlockOut res;
    res.end_pos = pos, res.out_len = 0, res.n_seg = 0, res.status = ZPAQGPU_OK;
    u8 *dst = A.out + blk.out_off;
    auto rd = [&](u64 at) -> i32 { return at < A.arc_len ? i32(arc[at]) : -1; };
    for (;;) {
        const i32 marker = rd(pos++);
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
        rec.out_off = blk.out_off + res.out_len;
        u64 produced = 0;
        bool first = true, truncated = false;
        for (;;) {
            const i32 b0 = rd(pos), b1 = rd(pos + 1), b2 = rd(pos + 2), b3 = rd(pos + 3);
            pos += 4;
            if (b0 < 0 || b1 < 0 || b2 < 0 || b3 < 0) { truncated = true; break; }
            u64 n = (u64(b0) << 24) | (u64(b1) << 16) | (u64(b2) << 8) | u64(b3);
            if (n == 0) break;
            if (first) {
                if (rd(pos++) < 0) { truncated = true; break; }
                --n, first = false;
                if (n == 0) continue;
            }
            const u64 avail = pos < A.arc_len ? A.arc_len - pos : 0;
            const u64 take = n < avail ? n : avail;
            const u64 base = res.out_len + produced;
            for (u64 q = lane; q < take; q += 32)
                if (base + q < blk.out_cap) dst[base + q] = arc[pos + q];
            produced += take, pos += take;
            if (take < n) { truncated = true; break; }
        }
        if (pos > A.arc_len) pos = A.arc_len;
        if (!truncated) {
            const i32 mk = rd(pos++);
            if (mk == 253) {
                rec.sha_off = pos;
                pos = min(pos + 20, A.arc_len);
            }
        }
        rec.out_len = produced;
        res.out_len += produced;
        if (lane == 0) {
            const u32 at = atomicAdd(A.seg_count, 1u);
            if (at < A.seg_cap) A.seg_recs[at] = rec;
        }
        res.n_seg++;
        if (truncated) { res.status = ZPAQGPU_E_FORMAT; break; }
    }
    if (pos > A.arc_len) pos = A.arc_len;
    res.end_pos = pos;
    if (lane == 0) A.results[bi] = res;
}

void launch_decode_store(const DecodeArgs &A, cudaStream_t s) {
    if (A.n_blocks <= 0) return;
    k_decode_store<<<(A.n_blocks + 3) / 4, 128, 0, s>>>(A);
}

__global__ void k_sha_compare(const u8 *arc, u64 arc_len, const DecSegRec *recs, int n, const u8 *digests,
                              i32 *ok) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u64 at = recs[j].sha_off;
    if (at == ~0ull) { ok[j] = -1; return; }
    i32 same = 1;
    for (int i = 0; i < 20; ++i) {
        // bytes past the end of the archive read as 0 (decompressor.v:610-616)
        const u32 stored = at + i < arc_len ? arc[at + i] : 0u;
        same &= (stored == digests[u64(j) * 20 + i]);
    }
    ok[j] = same;
}

void launch_sha_compare(const u8 *arc, u64 arc_len, const DecSegRec *recs, int n, const u8 *digests,
                        i32 *ok, cudaStream_t s) {
    if (n <= 0) return;
    k_sha_compare<<<(n + 127) / 128, 128, 0, s>>>(arc, arc_len, recs, n, digests, ok);
}

// k_gather_heads: the first `head` bytes of each block, for host-side header parsing when the
// archive only exists in device memory.
__global__ void k_gather_heads(const u8 *arc, const u64 *off, int n, u32 head, u8 *dst) {
    const int b = blockIdx.x;
    if (b >= n) return;
    const u64 lo = off[b], hi = off[b + 1];
    for (u32 i = threadIdx.x; i < head; i += blockDim.x) dst[u64(b) * head + i] = lo + i < hi ? arc[lo + i] : u8(0);
}
void launch_gather_heads(const u8 *arc, const u64 *off, int n, u32 head, u8 *dst, cudaStream_t s) {
    if (n <= 0) return;
    k_gather_heads<<<n, 128, 0, s>>>(arc, off, n, head, dst);
}

}  // namespace zg

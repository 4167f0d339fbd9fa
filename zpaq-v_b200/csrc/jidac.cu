// jidac.cu -- the jidac front end of libzpaqgpu: content-defined fragmentation, SHA-1 of every
// fragment, deduplication, and assembly of a journaling archive (c / d / h / i blocks) in the
// layout JidacArchive.create_archive writes (jidac.v:181-296).
//
// Stages, each its own kernel (all integer/byte work, HBM- or latency-bound, no tensor cores):
//   k_fragment_files   one thread per file walks its bytes with upstream zpaq's rolling hash
//                      (order-1 prediction table o1[256] per thread in shared memory)
//   k_frag_scan/_emit  per-file fragment counts -> one fragment list in file order
//   k_sha1_segments    SHA-1 per fragment (kernels_aux.cu, sha1.v:42-146)
//   k_dedup_insert/_resolve   open-addressing table keyed by the digest; the representative of a
//                      group of equal fragments is the smallest fragment index (deterministic ids)
//   k_frag_ids         stored flags, 1-based ids (jidac.v:153-163) and packed offsets, one scan
//   k_gather_fragments stored fragments back to back: the plaintext of the d blocks
// The d blocks then go through the same run_compress job as zpaqgpu_compress_blocks, the c/h/i
// blocks through the store-mode path of the same job (compressor.v:297-354).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "ctx.h"
#include "hostlogic.h"

using namespace zg;

namespace zg {
namespace {

constexpr int kFragThreads = 32;  // one warp per CTA: the prediction tables take 8 KiB of shared memory
constexpr u32 kEmpty = 0xFFFFFFFFu;

struct FragArgs {
    const u8 *in;
    const u64 *file_off;  // n_files + 1, relative to `in`
    const u32 *order;     // file handled by thread t: longest files first, so that the lanes of a
                          // warp walk files of similar length
    int n_files;
    int fragment;
    const u64 *cap_off;   // first record slot of each file
    u64 *ends;            // fragment end offsets, slot cap_off[f] + j
    u32 *count;           // fragments per file
};

// Upstream zpaq's cut rule (restated from memory of zpaq 7.15 `add`; not part of the reference, so
// parity is unpinned; the test suite checks this kernel against a separate CPU statement of the
// same rule): per fragment,
//   h = (h + c + 1) * (c == o1[c1] ? 314159265 : 271828182);  o1[c1] = c;  c1 = c
// and the fragment ends at EOF, at 8128<<fragment bytes, or when h < 2^(22-fragment) after at least
// 64<<fragment bytes.  h, c1 and o1 start from zero in every fragment, which makes the fragments of
// one file a serial chain; parallelism is across files.
__global__ void __launch_bounds__(kFragThreads) k_fragment_files(FragArgs A) {
    // o1[c1] of lane t lives in byte (c1 & 3) of word (c1 >> 2) * 32 + t: every lane has its own bank
    __shared__ u32 o1w[64 * kFragThreads];
    const int t = blockIdx.x * kFragThreads + threadIdx.x;
    if (t >= A.n_files) return;
    const u32 f = A.order[t];
    const u64 lo = A.file_off[f], hi = A.file_off[f + 1];
    u64 *ends = A.ends + A.cap_off[f];
    if (A.fragment < 0) {  // jidac.v:191-200: the whole file is one fragment, an empty file too
        ends[0] = hi;
        A.count[f] = 1;
        return;
    }
    const u64 minf = 64ull << A.fragment, maxf = 8128ull << A.fragment;
    const bool hashed = A.fragment <= 22;
    const u32 thresh = hashed ? (1u << (22 - A.fragment)) : 0u;
    u8 *mine = reinterpret_cast<u8 *>(o1w) + 4 * threadIdx.x;
    auto slot = [&](u32 c1) -> u8 * { return mine + (c1 >> 2) * (4 * kFragThreads) + (c1 & 3u); };
    const u8 *in = A.in;  // 16-byte aligned; the buffer has slack past the last byte
    u32 cnt = 0;
    u64 pos = lo;
    // One flat loop, one aligned 16-byte piece per step and lane (the next piece is requested before
    // this one is walked), so that the lanes of a warp -- files of similar length -- stay converged.
    // All bytes of a piece are hashed without looking at the cut condition: what the table and h hold
    // after a cut is thrown away by the reset that follows, and this keeps the table accesses and the
    // multiply chain free of the compare.
    uint4 nxt = *reinterpret_cast<const uint4 *>(in + (pos & ~15ull));
    for (int k = 0; k < 64; ++k) o1w[k * kFragThreads + threadIdx.x] = 0;
    u32 h = 0, c1 = 0;
    u64 stop = min(hi, pos + maxf), earliest = pos + minf;
    bool active = pos < hi;
    while (active) {
        const u64 base = pos & ~15ull;
        const uint4 v = nxt;
        nxt = *reinterpret_cast<const uint4 *>(in + base + 16);
        const u32 w4[4] = {v.x, v.y, v.z, v.w};
        const u32 j0 = u32(pos - base);
        const u32 j1 = stop - base < 16 ? u32(stop - base) : 16u;  // bytes of this piece inside the fragment limit
        const u32 je = earliest > base ? (earliest - base < 64 ? u32(earliest - base) : 64u) : 0u;
        u32 first_cut = 32;
#pragma unroll
        for (u32 j = 0; j < 16; ++j) {
            if (j >= j0) {
                const u32 c = (w4[j >> 2] >> (8 * (j & 3))) & 255u;
                u8 *at = slot(c1);
                const u32 pred = *at;
                h = (h + c + 1u) * (c == pred ? 314159265u : 271828182u);
                *at = u8(c);
                c1 = c;
                if (hashed && h < thresh && j + 1 >= je) first_cut = min(first_cut, j + 1);
            }
        }
        const u32 upto = min(first_cut, j1);  // first_cut beyond j1 lies past the size limit / end of file
        pos = base + upto;
        if (first_cut <= j1 || pos >= stop) {  // the fragment ends here
            ends[cnt++] = pos;
            for (int k = 0; k < 64; ++k) o1w[k * kFragThreads + threadIdx.x] = 0;
            h = 0, c1 = 0;
            stop = min(hi, pos + maxf), earliest = pos + minf;
            active = pos < hi;
            if ((pos & 15ull) != 0) nxt = v;  // the next fragment starts inside this piece
        }
    }
    A.count[f] = cnt;
}

// single CTA: base[] = exclusive prefix sum of count[], base[n] = total
__global__ void k_frag_scan(const u32 *count, int n, u32 *base) {
    __shared__ u32 partial[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int lo = min(n, t * per), hi = min(n, lo + per);
    u32 sum = 0;
    for (int i = lo; i < hi; ++i) sum += count[i];
    partial[t] = sum;
    __syncthreads();
    if (t == 0) {
        u32 run = 0;
        for (int i = 0; i < nt; ++i) {
            const u32 v = partial[i];
            partial[i] = run;
            run += v;
        }
        base[n] = run;
    }
    __syncthreads();
    u32 at = partial[t];
    for (int i = lo; i < hi; ++i) {
        base[i] = at;
        at += count[i];
    }
}

__global__ void k_frag_emit(const u64 *file_off, const u64 *cap_off, const u64 *ends, const u32 *count,
                            const u32 *base, int n_files, ShaJob *jobs, u32 *file_of) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_files) return;
    u64 start = file_off[f];
    const u64 *e = ends + cap_off[f];
    const u32 n = count[f], b = base[f];
    for (u32 j = 0; j < n; ++j) {
        jobs[b + j].off = start, jobs[b + j].len = e[j] - start;
        file_of[b + j] = u32(f);
        start = e[j];
    }
}

__device__ __forceinline__ bool same_fragment(const u8 *digests, const ShaJob *jobs, u32 a, u32 b) {
    if (jobs[a].len != jobs[b].len) return false;
    const u32 *x = reinterpret_cast<const u32 *>(digests + u64(a) * 20);
    const u32 *y = reinterpret_cast<const u32 *>(digests + u64(b) * 20);
    return x[0] == y[0] && x[1] == y[1] && x[2] == y[2] && x[3] == y[3] && x[4] == y[4];
}

// Every group of equal fragments ends up owning one table slot that holds its smallest index.
__global__ void k_dedup_insert(const u8 *digests, const ShaJob *jobs, u32 n, u32 *tab, u32 mask) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s = *reinterpret_cast<const u32 *>(digests + u64(i) * 20) & mask;
    for (;;) {
        const u32 cur = atomicCAS(&tab[s], kEmpty, i);
        if (cur == kEmpty) return;
        if (same_fragment(digests, jobs, cur, i)) {
            atomicMin(&tab[s], i);
            return;
        }
        s = (s + 1) & mask;
    }
}

__global__ void k_dedup_resolve(const u8 *digests, const ShaJob *jobs, u32 n, const u32 *tab, u32 mask,
                                u32 *rep) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s = *reinterpret_cast<const u32 *>(digests + u64(i) * 20) & mask;
    for (;;) {
        const u32 cur = tab[s];
        if (cur == kEmpty || same_fragment(digests, jobs, cur, i)) {  // kEmpty cannot happen after insert
            rep[i] = cur == kEmpty ? i : cur;
            return;
        }
        s = (s + 1) & mask;
    }
}

__global__ void k_iota(u32 *rep, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rep[i] = i;
}

// single CTA: stored[i] = (rep[i] == i); ids count stored fragments from 1 in list order
// (a.fragments.len after the push, jidac.v:153-163); pack_off = bytes of stored fragments before i;
// duplicates take the id of their representative.  totals = {n_stored, stored_bytes}.
__global__ void k_frag_ids(const ShaJob *jobs, const u32 *rep, u32 n, u32 *stored, u32 *id, u64 *pack_off,
                           u64 *totals) {
    __shared__ u32 pc[1024];
    __shared__ u64 pb[1024];
    const u32 t = threadIdx.x, nt = blockDim.x;
    const u32 per = (n + nt - 1) / nt;
    const u32 lo = min(n, t * per), hi = min(n, lo + per);
    u32 c = 0;
    u64 b = 0;
    for (u32 i = lo; i < hi; ++i)
        if (rep[i] == i) ++c, b += jobs[i].len;
    pc[t] = c, pb[t] = b;
    __syncthreads();
    if (t == 0) {
        u32 rc = 0;
        u64 rb = 0;
        for (u32 k = 0; k < nt; ++k) {
            const u32 vc = pc[k];
            const u64 vb = pb[k];
            pc[k] = rc, pb[k] = rb;
            rc += vc, rb += vb;
        }
        totals[0] = rc, totals[1] = rb;
    }
    __syncthreads();
    c = pc[t], b = pb[t];
    for (u32 i = lo; i < hi; ++i) {
        const bool s = rep[i] == i;
        stored[i] = s ? 1u : 0u;
        pack_off[i] = b;
        if (s) id[i] = ++c, b += jobs[i].len;
    }
    __syncthreads();
    __threadfence_block();
    for (u32 i = lo; i < hi; ++i)
        if (rep[i] != i) id[i] = id[rep[i]];
}

__global__ void k_gather_fragments(const u8 *in, const ShaJob *jobs, const u32 *stored, const u64 *pack_off,
                                   u8 *dst) {
    const u32 i = blockIdx.x;
    if (!stored[i]) return;
    const u8 *s = in + jobs[i].off;
    u8 *d = dst + pack_off[i];
    const u64 len = jobs[i].len;
    // 16-byte copies when source and destination are congruent mod 16, bytes otherwise
    if (((reinterpret_cast<uintptr_t>(s) ^ reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        const u64 head = min(len, u64((16 - (reinterpret_cast<uintptr_t>(s) & 15)) & 15));
        for (u64 k = threadIdx.x; k < head; k += blockDim.x) d[k] = s[k];
        const u64 vec = (len - head) / 16;
        const uint4 *sv = reinterpret_cast<const uint4 *>(s + head);
        uint4 *dv = reinterpret_cast<uint4 *>(d + head);
        for (u64 k = threadIdx.x; k < vec; k += blockDim.x) dv[k] = sv[k];
        for (u64 k = head + vec * 16 + threadIdx.x; k < len; k += blockDim.x) d[k] = s[k];
    } else {
        for (u64 k = threadIdx.x; k < len; k += blockDim.x) d[k] = s[k];
    }
}

struct CopyJob {
    u64 src, dst, len;
};
// one CTA per range: fragments of the d blocks -> their place in the restored files
__global__ void k_copy_ranges(const u8 *src_base, u8 *dst_base, const CopyJob *jobs) {
    const CopyJob j = jobs[blockIdx.x];
    const u8 *s = src_base + j.src;
    u8 *d = dst_base + j.dst;
    if (((reinterpret_cast<uintptr_t>(s) ^ reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        const u64 head = min(j.len, u64((16 - (reinterpret_cast<uintptr_t>(s) & 15)) & 15));
        for (u64 k = threadIdx.x; k < head; k += blockDim.x) d[k] = s[k];
        const u64 vec = (j.len - head) / 16;
        const uint4 *sv = reinterpret_cast<const uint4 *>(s + head);
        uint4 *dv = reinterpret_cast<uint4 *>(d + head);
        for (u64 k = threadIdx.x; k < vec; k += blockDim.x) dv[k] = sv[k];
        for (u64 k = head + vec * 16 + threadIdx.x; k < j.len; k += blockDim.x) d[k] = s[k];
    } else {
        for (u64 k = threadIdx.x; k < j.len; k += blockDim.x) d[k] = s[k];
    }
}

// ---- host side ---------------------------------------------------------------------------

struct Front {
    int n_frags = 0, n_stored = 0;
    u64 stored_bytes = 0, total_in = 0;
    std::vector<ShaJob> jobs;  // ranges relative to the first input byte
    std::vector<u32> file_of, id, stored;
    std::vector<u64> pack_off;
    std::vector<u8> digests;
    const u8 *d_plain = nullptr;  // device: stored fragments back to back in id order
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    Timer() { cudaEventCreate(&a), cudaEventCreate(&b); }
    ~Timer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    void stop(cudaStream_t s) { cudaEventRecord(b, s); }
    float ms() {
        cudaEventSynchronize(b);
        return elapsed(a, b);
    }
};

int jidac_front(zpaqgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off, int n_files, int fragment, int dedup,
                bool gather, Front &R) {
    Range nvtx_front("zpaqgpu:jidac_front");
    zpaqgpu_jidac_stats &S = ctx->jd_stats;
    S = zpaqgpu_jidac_stats{};
    S.n_files = n_files;
    // -1: one fragment per file; 0..22 rolling hash; above: fixed size.  Beyond 40 the 64-bit fragment
    // limits (64 << fragment, 8128 << fragment) would wrap.
    if (fragment < -1 || fragment > 40) {
        ctx->err = "fragment must be in -1..40";
        return ZPAQGPU_E_ARG;
    }
    if (n_files == 0) return ZPAQGPU_OK;
    cudaStream_t st = ctx->stream;
    const u64 base = in_off[0];
    R.total_in = in_off[n_files] - base;
    S.input_bytes = R.total_in;
    if (R.total_in && !in) return ZPAQGPU_E_ARG;
    int rc;
    if ((rc = ensure(ctx, ctx->jd_in, R.total_in + 64))) return rc;  // k_fragment_files reads whole 16-byte pieces
    Timer t_h2d, t_frag, t_sha, t_dedup, t_gather;
    t_h2d.start(st);
    if (R.total_in) CK(cudaMemcpyAsync(ctx->jd_in.p, in + base, R.total_in, cudaMemcpyHostToDevice, st));
    t_h2d.stop(st);

    // per-file tables: offsets, processing order (longest first), record capacity
    const size_t nf = size_t(n_files);
    std::vector<u64> file_off(nf + 1), cap_off(nf + 1);
    std::vector<u32> order(nf);
    const u64 minf = fragment < 0 ? 0 : (64ull << std::min(fragment, 40));
    u64 cap_total = 0;
    for (size_t f = 0; f <= nf; ++f) {
        if (f && in_off[f] < in_off[f - 1]) return ZPAQGPU_E_ARG;
        file_off[f] = in_off[f] - base;
    }
    for (size_t f = 0; f < nf; ++f) {
        cap_off[f] = cap_total;
        const u64 len = file_off[f + 1] - file_off[f];
        cap_total += fragment < 0 ? 1 : len / minf + 1;
    }
    cap_off[nf] = cap_total;
    if (cap_total >= 0xFFFFFFF0ull) {
        ctx->err = "too many fragments for one call";
        return ZPAQGPU_E_ARG;
    }
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) {
        return file_off[a + 1] - file_off[a] > file_off[b + 1] - file_off[b];
    });
    // device layout of the fragment tables
    const size_t ct = std::max<size_t>(size_t(cap_total), 1);
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t o = at; at = align_up(at + bytes, 16); return o; };
    const size_t o_foff = take(8 * (nf + 1)), o_coff = take(8 * (nf + 1)), o_order = take(4 * nf);
    const size_t o_count = take(4 * nf), o_base = take(4 * (nf + 1));
    const size_t o_ends = take(8 * ct), o_jobs = take(sizeof(ShaJob) * ct), o_file = take(4 * ct);
    const size_t o_rep = take(4 * ct), o_id = take(4 * ct), o_stored = take(4 * ct), o_pack = take(8 * ct);
    const size_t o_dig = take(20 * ct + 16), o_tot = take(16);
    if ((rc = ensure(ctx, ctx->jd_frag, at))) return rc;
    u8 *fb = static_cast<u8 *>(ctx->jd_frag.p);
    CK(cudaMemcpyAsync(fb + o_foff, file_off.data(), 8 * (nf + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(fb + o_coff, cap_off.data(), 8 * (nf + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(fb + o_order, order.data(), 4 * nf, cudaMemcpyHostToDevice, st));
    const u8 *d_in = static_cast<const u8 *>(ctx->jd_in.p);
    ShaJob *d_jobs = reinterpret_cast<ShaJob *>(fb + o_jobs);
    u32 *d_rep = reinterpret_cast<u32 *>(fb + o_rep), *d_id = reinterpret_cast<u32 *>(fb + o_id);
    u32 *d_stored = reinterpret_cast<u32 *>(fb + o_stored), *d_base = reinterpret_cast<u32 *>(fb + o_base);
    u64 *d_pack = reinterpret_cast<u64 *>(fb + o_pack);
    u8 *d_dig = fb + o_dig;

    t_frag.start(st);
    FragArgs fa;
    fa.in = d_in, fa.file_off = reinterpret_cast<const u64 *>(fb + o_foff);
    fa.order = reinterpret_cast<const u32 *>(fb + o_order), fa.n_files = n_files, fa.fragment = fragment;
    fa.cap_off = reinterpret_cast<const u64 *>(fb + o_coff), fa.ends = reinterpret_cast<u64 *>(fb + o_ends);
    fa.count = reinterpret_cast<u32 *>(fb + o_count);
    k_fragment_files<<<(n_files + kFragThreads - 1) / kFragThreads, kFragThreads, 0, st>>>(fa);
    k_frag_scan<<<1, 1024, 0, st>>>(fa.count, n_files, d_base);
    k_frag_emit<<<(n_files + 127) / 128, 128, 0, st>>>(fa.file_off, fa.cap_off, fa.ends, fa.count, d_base, n_files,
                                                       d_jobs, reinterpret_cast<u32 *>(fb + o_file));
    t_frag.stop(st);
    CK(cudaGetLastError());
    u32 n_frags = 0;
    CK(cudaMemcpyAsync(&n_frags, d_base + n_files, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    S.launches += 3;
    R.n_frags = int(n_frags);
    S.n_fragments = R.n_frags;
    S.fragment_ms = t_frag.ms();
    S.h2d_ms = t_h2d.ms();
    if (n_frags == 0) return ZPAQGPU_OK;

    t_sha.start(st);
    launch_sha1(d_in, d_jobs, int(n_frags), d_dig, st);
    t_sha.stop(st);
    S.launches += 1;

    t_dedup.start(st);
    const u32 grid = (n_frags + 255) / 256;
    if (dedup) {
        u32 slots = 64;
        while (slots < 2 * n_frags) slots <<= 1;
        if ((rc = ensure(ctx, ctx->jd_tab, 4 * size_t(slots)))) return rc;
        CK(cudaMemsetAsync(ctx->jd_tab.p, 0xFF, 4 * size_t(slots), st));
        u32 *tab = static_cast<u32 *>(ctx->jd_tab.p);
        k_dedup_insert<<<grid, 256, 0, st>>>(d_dig, d_jobs, n_frags, tab, slots - 1);
        k_dedup_resolve<<<grid, 256, 0, st>>>(d_dig, d_jobs, n_frags, tab, slots - 1, d_rep);
        S.launches += 3;
    } else {
        k_iota<<<grid, 256, 0, st>>>(d_rep, n_frags);
        S.launches += 1;
    }
    k_frag_ids<<<1, 1024, 0, st>>>(d_jobs, d_rep, n_frags, d_stored, d_id, d_pack,
                                   reinterpret_cast<u64 *>(fb + o_tot));
    t_dedup.stop(st);
    S.launches += 1;
    CK(cudaGetLastError());

    R.jobs.resize(n_frags), R.file_of.resize(n_frags), R.id.resize(n_frags), R.stored.resize(n_frags);
    R.pack_off.resize(n_frags), R.digests.resize(20 * size_t(n_frags));
    u64 totals[2] = {0, 0};
    CK(cudaMemcpyAsync(R.jobs.data(), d_jobs, sizeof(ShaJob) * n_frags, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(R.file_of.data(), fb + o_file, 4 * size_t(n_frags), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(R.id.data(), d_id, 4 * size_t(n_frags), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(R.stored.data(), d_stored, 4 * size_t(n_frags), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(R.pack_off.data(), d_pack, 8 * size_t(n_frags), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(R.digests.data(), d_dig, 20 * size_t(n_frags), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(totals, fb + o_tot, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    R.n_stored = int(totals[0]), R.stored_bytes = totals[1];
    S.n_stored = R.n_stored, S.stored_bytes = R.stored_bytes;
    S.sha1_ms = t_sha.ms(), S.dedup_ms = t_dedup.ms();

    R.d_plain = d_in;  // without duplicates the fragments tile the input: pack_off == off
    if (gather && R.n_stored != R.n_frags) {
        if ((rc = ensure(ctx, ctx->jd_packed, std::max<u64>(R.stored_bytes, 16)))) return rc;
        t_gather.start(st);
        k_gather_fragments<<<n_frags, 256, 0, st>>>(d_in, d_jobs, d_stored, d_pack, static_cast<u8 *>(ctx->jd_packed.p));
        t_gather.stop(st);
        CK(cudaGetLastError());
        S.launches += 1;
        S.gather_ms = t_gather.ms();
        R.d_plain = static_cast<const u8 *>(ctx->jd_packed.p);
    }
    return ZPAQGPU_OK;
}

struct DBlock {
    u32 first_id, n;
    u64 off, len;        // inside the packed plaintext
    std::vector<u32> frags;  // fragment list indices
};

}  // namespace
}  // namespace zg

extern "C" {

int zpaqgpu_jidac_last_stats(const zpaqgpu_ctx *ctx, zpaqgpu_jidac_stats *out) {
    if (!ctx || !out) return ZPAQGPU_E_ARG;
    *out = ctx->jd_stats;
    return ZPAQGPU_OK;
}

int zpaqgpu_jidac_fragment(zpaqgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off, int n_files, int fragment,
                           int dedup, zpaqgpu_fragment *frags, int cap, int *n_frags, int *n_stored) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || n_files < 0 || (n_files > 0 && !in_off) || !n_frags) return ZPAQGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Front R;
    const int rc = jidac_front(ctx, in, in_off, n_files, fragment, dedup, false, R);
    if (rc) return rc;
    *n_frags = R.n_frags;
    if (n_stored) *n_stored = R.n_stored;
    if (R.n_frags > cap) return ZPAQGPU_E_NOSPACE;
    if (R.n_frags && !frags) return ZPAQGPU_E_ARG;
    const u64 base = n_files ? in_off[0] : 0;
    for (int i = 0; i < R.n_frags; ++i) {
        zpaqgpu_fragment &f = frags[i];
        f.off = R.jobs[size_t(i)].off + base, f.len = R.jobs[size_t(i)].len;
        f.file = R.file_of[size_t(i)], f.id = R.id[size_t(i)], f.stored = R.stored[size_t(i)];
        std::memcpy(f.sha1, R.digests.data() + 20 * size_t(i), 20);
    }
    return ZPAQGPU_OK;
    });
}

int zpaqgpu_jidac_extract(zpaqgpu_ctx *ctx, const uint8_t *arc, uint64_t len, uint8_t *out, uint64_t out_cap,
                          uint64_t *out_need, zpaqgpu_jidac_file *files, int files_cap, int *n_files, char *names,
                          uint64_t names_cap, uint64_t *names_need) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || (len && !arc)) return ZPAQGPU_E_ARG;
    if (out_need) *out_need = 0;
    if (n_files) *n_files = 0;
    if (names_need) *names_need = 0;
    if (len == 0) return ZPAQGPU_OK;
    std::vector<DecodedSeg> segs;
    const u8 *d_plain = nullptr;
    int status = ZPAQGPU_OK;
    u64 total = 0;
    int rc = decode_archive_dev(ctx, arc, len, segs, &d_plain, &status, &total);
    if (rc) return rc;
    if (status != ZPAQGPU_OK) return status;
    cudaStream_t st = ctx->stream;
    ctx->jd_stats = zpaqgpu_jidac_stats{};
    ctx->jd_stats.codec_ms = ctx->stats.codec_ms, ctx->jd_stats.h2d_ms = ctx->stats.h2d_ms;
    ctx->jd_stats.launches = ctx->stats.launches;

    // journaling blocks by kind: name = jDC<date14><kind><num10> (jidac.v:47-49)
    struct JSeg { char kind; u32 num; const DecodedSeg *s; };
    std::vector<JSeg> js;
    for (const DecodedSeg &d : segs) {
        const u64 at = d.seg.name_off;
        if (at + 28 >= len || std::memcmp(arc + at, "jDC", 3) != 0 || arc[at + 28] != 0) continue;
        bool ok = true;
        u64 num = 0;
        for (int k = 3; k < 17 && ok; ++k) ok = arc[at + k] >= '0' && arc[at + k] <= '9';
        for (int k = 18; k < 28 && ok; ++k) ok = arc[at + k] >= '0' && arc[at + k] <= '9', num = num * 10 + (arc[at + k] - '0');
        const char kind = char(arc[at + 17]);
        if (!ok || (kind != 'c' && kind != 'd' && kind != 'h' && kind != 'i')) continue;
        if (d.seg.sha1_ok == 0) {
            ctx->err = "SHA-1 of a journaling block does not match";
            return ZPAQGPU_E_FORMAT;
        }
        js.push_back(JSeg{kind, u32(num), &d});
    }
    // the index blocks are small: bring their plaintext to the host
    u64 idx_bytes = 0;
    for (const JSeg &j : js)
        if (j.kind == 'h' || j.kind == 'i') idx_bytes += j.s->seg.out_len;
    std::vector<u8> idx(static_cast<size_t>(idx_bytes) + 1);
    {
        u64 at = 0;
        for (const JSeg &j : js)
            if (j.kind == 'h' || j.kind == 'i') {
                if (j.s->seg.out_len)
                    CK(cudaMemcpyAsync(idx.data() + at, d_plain + j.s->src, j.s->seg.out_len, cudaMemcpyDeviceToHost, st));
                at += j.s->seg.out_len;
            }
        CK(cudaStreamSynchronize(st));
    }
    auto le = [](const u8 *p, int n) { u64 v = 0; for (int i = 0; i < n; ++i) v |= u64(p[i]) << (8 * i); return v; };
    // d blocks by their first fragment id; fragments from the h tables
    struct Frag { u64 src; u32 len; u8 sha1[20]; bool known = false; };
    std::vector<Frag> frags(1);  // ids are 1-based
    // an id can never exceed the number of table entries in the archive: a 10-digit block number from
    // a damaged or crafted archive must not size an allocation
    u64 h_entries = 0;
    for (const JSeg &j : js)
        if (j.kind == 'h' && j.s->seg.out_len >= 4) h_entries += (j.s->seg.out_len - 4) / 24;
    std::vector<std::pair<u32, const DecodedSeg *>> dsegs;
    for (const JSeg &j : js)
        if (j.kind == 'd') dsegs.emplace_back(j.num, j.s);
    struct FileRec { std::string name; i64 date; std::vector<u32> ptr; bool live; };
    std::vector<FileRec> recs;
    {
        u64 at = 0;
        for (const JSeg &j : js) {
            if (j.kind != 'h' && j.kind != 'i') continue;
            const u8 *p = idx.data() + at;
            const u64 n = j.s->seg.out_len;
            at += n;
            if (j.kind == 'h') {
                const DecodedSeg *dblk = nullptr;
                for (const auto &d : dsegs)
                    if (d.first == j.num) dblk = d.second;
                if (!dblk || n < 4 || (n - 4) % 24 != 0) {
                    ctx->err = "fragment table without its data block";
                    return ZPAQGPU_E_FORMAT;
                }
                u64 off = 0;
                u32 id = j.num;
                for (u64 q = 4; q + 24 <= n; q += 24, ++id) {
                    if (id == 0 || u64(id) > h_entries) {
                        ctx->err = "fragment id outside the tables of the archive";
                        return ZPAQGPU_E_FORMAT;
                    }
                    if (frags.size() <= id) frags.resize(size_t(id) + 1);
                    Frag &f = frags[id];
                    std::memcpy(f.sha1, p + q, 20);
                    f.len = u32(le(p + q + 20, 4));
                    f.src = dblk->src + off, f.known = true;
                    off += f.len;
                }
                if (off != dblk->seg.out_len) {
                    ctx->err = "fragment sizes do not add up to the data block";
                    return ZPAQGPU_E_FORMAT;
                }
            } else {
                u64 q = 0;
                while (q + 8 <= n) {
                    FileRec r;
                    r.date = i64(le(p + q, 8)), q += 8;
                    u64 e = q;
                    while (e < n && p[e]) ++e;
                    if (e >= n) return ctx->err = "unterminated name in the index", ZPAQGPU_E_FORMAT;
                    r.name.assign(reinterpret_cast<const char *>(p + q), size_t(e - q));
                    q = e + 1;
                    r.live = r.date != 0;
                    if (r.live) {
                        if (q + 4 > n) return ctx->err = "truncated index entry", ZPAQGPU_E_FORMAT;
                        const u64 na = le(p + q, 4);
                        q += 4 + na;
                        if (q + 4 > n) return ctx->err = "truncated index entry", ZPAQGPU_E_FORMAT;
                        const u64 ni = le(p + q, 4);
                        q += 4;
                        if (q + 4 * ni > n) return ctx->err = "truncated index entry", ZPAQGPU_E_FORMAT;
                        for (u64 k = 0; k < ni; ++k) r.ptr.push_back(u32(le(p + q + 4 * k, 4)));
                        q += 4 * ni;
                    }
                    // a later entry of the same name replaces the earlier one; date 0 removes it
                    for (FileRec &old : recs)
                        if (old.live && old.name == r.name) old.live = false;
                    recs.push_back(std::move(r));
                }
            }
        }
    }
    // output layout and copy list
    std::vector<CopyJob> jobs;
    std::vector<const FileRec *> live;
    u64 out_total = 0, name_total = 0;
    for (const FileRec &r : recs) {
        if (!r.live) continue;
        live.push_back(&r);
        name_total += r.name.size() + 1;
        for (u32 id : r.ptr) {
            if (id == 0 || id >= frags.size() || !frags[id].known)
                return ctx->err = "index refers to an unknown fragment", ZPAQGPU_E_FORMAT;
            if (frags[id].len) jobs.push_back(CopyJob{frags[id].src, out_total, frags[id].len});
            out_total += frags[id].len;
        }
    }
    if (out_need) *out_need = out_total;
    if (n_files) *n_files = int(live.size());
    if (names_need) *names_need = name_total;
    if (out_total > out_cap || int(live.size()) > files_cap || name_total > names_cap) return ZPAQGPU_E_NOSPACE;
    if ((out_total && !out) || (!live.empty() && (!files || !names))) return ZPAQGPU_E_ARG;
    // SHA-1 of every known fragment against its table entry
    const size_t nfr = frags.size();
    std::vector<ShaJob> sj;
    std::vector<u32> sj_id;
    for (size_t id = 1; id < nfr; ++id)
        if (frags[id].known) sj.push_back(ShaJob{frags[id].src, frags[id].len}), sj_id.push_back(u32(id));
    std::vector<u8> dig(20 * sj.size() + 1);
    std::vector<char> frag_ok(nfr, 0);
    Timer t_sha, t_gather, t_d2h;
    if (!sj.empty()) {
        if ((rc = ensure(ctx, ctx->jd_frag, sizeof(ShaJob) * sj.size() + 20 * sj.size() + 64))) return rc;
        u8 *fb = static_cast<u8 *>(ctx->jd_frag.p);
        const size_t o_dig = align_up(sizeof(ShaJob) * sj.size(), 16);
        CK(cudaMemcpyAsync(fb, sj.data(), sizeof(ShaJob) * sj.size(), cudaMemcpyHostToDevice, st));
        t_sha.start(st);
        launch_sha1(d_plain, reinterpret_cast<const ShaJob *>(fb), int(sj.size()), fb + o_dig, st);
        t_sha.stop(st);
        CK(cudaMemcpyAsync(dig.data(), fb + o_dig, 20 * sj.size(), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->jd_stats.sha1_ms = t_sha.ms();
        ctx->jd_stats.launches += 1;
        for (size_t k = 0; k < sj.size(); ++k)
            frag_ok[sj_id[k]] = std::memcmp(dig.data() + 20 * k, frags[sj_id[k]].sha1, 20) == 0;
    }
    // gather on the device, one copy out
    if (out_total) {
        if ((rc = ensure(ctx, ctx->jd_packed, out_total))) return rc;
        if ((rc = ensure(ctx, ctx->jd_small, sizeof(CopyJob) * jobs.size() + 16))) return rc;
        if (!jobs.empty()) {
            CK(cudaMemcpyAsync(ctx->jd_small.p, jobs.data(), sizeof(CopyJob) * jobs.size(), cudaMemcpyHostToDevice, st));
            t_gather.start(st);
            k_copy_ranges<<<unsigned(jobs.size()), 256, 0, st>>>(d_plain, static_cast<u8 *>(ctx->jd_packed.p),
                                                                static_cast<const CopyJob *>(ctx->jd_small.p));
            t_gather.stop(st);
            CK(cudaGetLastError());
            ctx->jd_stats.launches += 1;
        }
        t_d2h.start(st);
        CK(cudaMemcpyAsync(out, ctx->jd_packed.p, out_total, cudaMemcpyDeviceToHost, st));
        t_d2h.stop(st);
        CK(cudaStreamSynchronize(st));
        if (!jobs.empty()) ctx->jd_stats.gather_ms = t_gather.ms();
        ctx->jd_stats.d2h_ms = t_d2h.ms();
    }
    u64 at = 0, nat = 0;
    for (size_t k = 0; k < live.size(); ++k) {
        const FileRec &r = *live[k];
        zpaqgpu_jidac_file &f = files[k];
        f.name_off = nat, f.out_off = at, f.date = r.date, f.n_fragments = int(r.ptr.size());
        std::memcpy(names + nat, r.name.c_str(), r.name.size() + 1);
        nat += r.name.size() + 1;
        u64 flen = 0;
        int ok = 1;
        for (u32 id : r.ptr) flen += frags[id].len, ok &= frag_ok[id] ? 1 : 0;
        f.out_len = flen, f.sha1_ok = ok;
        at += flen;
    }
    ctx->jd_stats.n_files = int(live.size()), ctx->jd_stats.n_fragments = int(sj.size());
    ctx->jd_stats.input_bytes = len, ctx->jd_stats.stored_bytes = total, ctx->jd_stats.archive_bytes = len;
    return ZPAQGPU_OK;
    });
}

int zpaqgpu_jidac_add(zpaqgpu_ctx *ctx, const zpaqgpu_jidac_opts *opts, const char *const *names, const uint8_t *in,
                      const uint64_t *in_off, int n_files, uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                      uint64_t *out_need) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || !opts || n_files < 0 || (n_files > 0 && (!in_off || !names))) return ZPAQGPU_E_ARG;
    if (opts->level < 0 || opts->level > 5) return ZPAQGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Front R;
    int rc = jidac_front(ctx, in, in_off, n_files, opts->fragment, opts->dedup, true, R);
    if (rc) return rc;
    zpaqgpu_jidac_stats S = ctx->jd_stats;  // run_compress below does not touch jd_stats

    // ---- d blocks: stored fragments in id order, packed up to block_bytes (jidac.v:186-214) ----
    std::vector<DBlock> dblocks;
    for (int i = 0; i < R.n_frags; ++i) {
        if (!R.stored[size_t(i)]) continue;
        const u64 len = R.jobs[size_t(i)].len;
        const bool fresh = dblocks.empty() || opts->block_bytes == 0 || dblocks.back().len + len > opts->block_bytes;
        if (fresh) dblocks.push_back(DBlock{R.id[size_t(i)], 0, R.pack_off[size_t(i)], 0, {}});
        DBlock &b = dblocks.back();
        b.n++, b.len += len, b.frags.push_back(u32(i));
    }
    const int nd = int(dblocks.size());
    S.n_dblocks = nd;
    std::vector<u64> d_off(size_t(nd) + 1, 0);
    u64 total_d = 0;
    if (nd) {
        const std::vector<uint8_t> h = level_header(opts->level);
        Model m;
        if ((rc = model_from_level_layout(h.data(), int(h.size()), m))) return ctx->err = m.error, rc;
        CompressJob job;
        job.model = &m;
        job.blocks.resize(size_t(nd)), job.segs.resize(size_t(nd));
        std::vector<std::string> nm(static_cast<size_t>(nd)), cm(static_cast<size_t>(nd));
        u64 worst = 0;
        for (int b = 0; b < nd; ++b) {
            const DBlock &d = dblocks[size_t(b)];
            nm[size_t(b)] = jidac_name(opts->date, 'd', d.first_id);
            cm[size_t(b)] = jidac_comment(d.len);
            job.blocks[size_t(b)] = EncBlock{u32(b), 1};
            SegSpec &sp = job.segs[size_t(b)];
            sp.name = nm[size_t(b)].c_str(), sp.comment = cm[size_t(b)].c_str();
            sp.in_off = d.off, sp.in_len = d.len, sp.called = true;  // jidac.v:112 always calls compress()
            worst += d.len + d.len / 2 + 2048;
        }
        if ((rc = ensure(ctx, ctx->out, worst))) return rc;
        if ((rc = ensure(ctx, ctx->out_off, 8 * size_t(nd + 1)))) return rc;
        job.d_in = R.d_plain;
        job.d_out = static_cast<u8 *>(ctx->out.p), job.out_cap = ctx->out.cap;
        job.d_out_off = static_cast<u64 *>(ctx->out_off.p);
        if ((rc = run_compress(ctx, job))) return rc;
        S.codec_ms += ctx->stats.codec_ms, S.pack_ms += ctx->stats.pack_ms, S.sha1_ms += ctx->stats.sha1_ms;
        S.launches += ctx->stats.launches;
        if (!job.fits) {
            if ((rc = ensure(ctx, ctx->out, job.total))) return rc;
            job.d_out = static_cast<u8 *>(ctx->out.p), job.out_cap = ctx->out.cap;
            if ((rc = run_compress(ctx, job))) return rc;
            S.codec_ms += ctx->stats.codec_ms, S.launches += ctx->stats.launches;
        }
        total_d = job.total;
        CK(cudaMemcpyAsync(d_off.data(), ctx->out_off.p, 8 * size_t(nd + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }

    // ---- c, h and i blocks, all store mode (jidac.v:67-91, :216-295) ----
    std::vector<u8> small;                 // their plaintext back to back
    std::vector<u64> s_off{0};
    std::vector<std::string> s_name, s_comment;
    auto close_block = [&](char type, u32 num) {
        s_name.push_back(jidac_name(opts->date, type, num));
        s_comment.push_back(jidac_comment(small.size() - s_off.back()));
        s_off.push_back(small.size());
    };
    put_le(small, total_d, 8);             // c block: bytes of all d blocks (jidac.v:217-219)
    close_block('c', u32(R.n_stored) + 1);
    for (int b = 0; b < nd; ++b) {         // h blocks: bsize[4] (sha1[20] usize[4])... (jidac.v:229-259)
        const DBlock &d = dblocks[size_t(b)];
        put_le(small, u32(d_off[size_t(b) + 1] - d_off[size_t(b)]), 4);
        for (u32 i : d.frags) {
            small.insert(small.end(), R.digests.begin() + 20 * size_t(i), R.digests.begin() + 20 * size_t(i) + 20);
            put_le(small, u32(R.jobs[i].len), 4);
        }
        close_block('h', d.first_id);
    }
    {                                      // i block: date[8] filename 0 na[4] ni[4] ptr[ni][4] (jidac.v:262-295)
        size_t fi = 0;
        for (int f = 0; f < n_files; ++f) {
            put_le(small, u64(opts->date), 8);
            for (const char *c = names[f] ? names[f] : ""; *c; ++c) small.push_back(u8(*c));
            small.push_back(0);
            size_t fe = fi;
            while (fe < size_t(R.n_frags) && R.file_of[fe] == u32(f)) ++fe;
            if (opts->date != 0) {
                put_le(small, 0, 4);
                put_le(small, u32(fe - fi), 4);
                for (size_t k = fi; k < fe; ++k) put_le(small, R.id[k], 4);
            }
            fi = fe;
        }
        if (small.size() > s_off.back()) close_block('i', 1);
    }
    const int ns = int(s_name.size());
    std::vector<u64> o2(size_t(ns) + 1, 0);
    u64 total_s = 0;
    {
        const std::vector<uint8_t> h = level_header(0);
        Model m0;
        if ((rc = model_from_level_layout(h.data(), int(h.size()), m0))) return ctx->err = m0.error, rc;
        if ((rc = ensure(ctx, ctx->jd_small, small.size() + 16))) return rc;
        CK(cudaMemcpyAsync(ctx->jd_small.p, small.data(), small.size(), cudaMemcpyHostToDevice, st));
        CompressJob job;
        job.model = &m0;
        job.blocks.resize(size_t(ns)), job.segs.resize(size_t(ns));
        for (int b = 0; b < ns; ++b) {
            job.blocks[size_t(b)] = EncBlock{u32(b), 1};
            SegSpec &sp = job.segs[size_t(b)];
            sp.name = s_name[size_t(b)].c_str(), sp.comment = s_comment[size_t(b)].c_str();
            sp.in_off = s_off[size_t(b)], sp.in_len = s_off[size_t(b) + 1] - s_off[size_t(b)], sp.called = true;
        }
        const u64 cap2 = small.size() + small.size() / 4096 + 256 * u64(ns) + 4096;
        if ((rc = ensure(ctx, ctx->jd_out2, cap2))) return rc;
        if ((rc = ensure(ctx, ctx->jd_off2, 8 * size_t(ns + 1)))) return rc;
        job.d_in = static_cast<const u8 *>(ctx->jd_small.p);
        job.d_out = static_cast<u8 *>(ctx->jd_out2.p), job.out_cap = ctx->jd_out2.cap;
        job.d_out_off = static_cast<u64 *>(ctx->jd_off2.p);
        if ((rc = run_compress(ctx, job))) return rc;
        S.launches += ctx->stats.launches, S.pack_ms += ctx->stats.pack_ms;
        if (!job.fits) {
            ctx->err = "index blocks larger than their bound";
            return ZPAQGPU_E_CUDA;
        }
        total_s = job.total;
        CK(cudaMemcpyAsync(o2.data(), ctx->jd_off2.p, 8 * size_t(ns + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    const u64 total = total_d + total_s;
    S.archive_bytes = total;
    if (out_len) *out_len = total;
    if (out_need) *out_need = total;
    ctx->jd_stats = S;
    if (total > out_cap) return ZPAQGPU_E_NOSPACE;
    if (!out) return ZPAQGPU_E_ARG;
    // archive order: c block, d blocks, h blocks, i block (jidac.v:221-295)
    Timer t_d2h;
    t_d2h.start(st);
    const u64 c_len = o2[1];
    const u8 *s2 = static_cast<const u8 *>(ctx->jd_out2.p);
    CK(cudaMemcpyAsync(out, s2, c_len, cudaMemcpyDeviceToHost, st));
    if (total_d) CK(cudaMemcpyAsync(out + c_len, ctx->out.p, total_d, cudaMemcpyDeviceToHost, st));
    if (total_s > c_len)
        CK(cudaMemcpyAsync(out + c_len + total_d, s2 + c_len, total_s - c_len, cudaMemcpyDeviceToHost, st));
    t_d2h.stop(st);
    CK(cudaStreamSynchronize(st));
    ctx->jd_stats.d2h_ms = t_d2h.ms();
    return ZPAQGPU_OK;
    });
}

}  // extern "C"

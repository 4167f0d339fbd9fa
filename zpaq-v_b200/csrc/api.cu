// api.cu -- the C ABI of libzpaqgpu (include/zpaqgpu.h): context, device-buffer management,
// wave scheduling over the table workspace and the host side of block framing.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ctx.h"
#include "hostlogic.h"

using namespace zg;

namespace zg {

int ensure(zpaqgpu_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return ZPAQGPU_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(b.p));
        b.p = nullptr, b.cap = 0;
    }
    // small buffers get 12.5 % head room against regrowth; the multi-gigabyte ones are sized exactly
    size_t want = bytes + (bytes < (size_t(1) << 30) ? bytes / 8 : 0) + 4096;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        cudaGetLastError();
        ctx->err = "cudaMalloc of " + std::to_string(bytes) + " bytes failed (" + std::to_string(free_b) + " of " +
                   std::to_string(total_b) + " bytes free)";
        return ZPAQGPU_E_NOMEM;
    }
    b.cap = want;
    return ZPAQGPU_OK;
}

int ensure_pinned(zpaqgpu_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return ZPAQGPU_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr, ctx->pinned_cap = 0;
    CK(cudaMallocHost(&ctx->pinned, bytes + 4096));
    ctx->pinned_cap = bytes + 4096;
    return ZPAQGPU_OK;
}

u64 align_up(u64 v, u64 a) { return (v + a - 1) / a * a; }

// The two table buffers of a wave.  plan_tables budgets with both of them counted as free, so when either
// has to grow both are released first: a large dense workspace kept from an earlier call must not sit
// beside a new page pool (or the other way round).
int ensure_tables(zpaqgpu_ctx *ctx, size_t ws_bytes, size_t pool_bytes) {
    if (ws_bytes <= ctx->workspace.cap && pool_bytes <= ctx->pool.cap) return ZPAQGPU_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    for (DevBuf *b : {&ctx->workspace, &ctx->pool})
        if (b->p) {
            CK(cudaFree(b->p));
            b->p = nullptr, b->cap = 0;
        }
    int rc = ensure(ctx, ctx->workspace, std::max<size_t>(ws_bytes, 256));
    if (rc == ZPAQGPU_OK && pool_bytes) rc = ensure(ctx, ctx->pool, pool_bytes);
    return rc;
}

// The model on the device: header bytes, both component layouts, their fill lists, the images.
struct ModelOnDev {
    ModelDev dense, paged;
    const FillRegion *fills_dense = nullptr, *fills_paged = nullptr;
    int n_fills_dense = 0, n_fills_paged = 0;
    const u32 *image = nullptr;
};

int upload_model(zpaqgpu_ctx *ctx, const Model &m, ModelOnDev &out) {
    auto pad = [](size_t v) { return align_up(v, 16); };
    const size_t n_c = std::max<size_t>(1, m.comps.size());
    const size_t o_hdr = 0;
    const size_t o_cd = pad(o_hdr + m.header.size() + 1);
    const size_t o_cp = pad(o_cd + sizeof(CompDesc) * n_c);
    const size_t o_fd = pad(o_cp + sizeof(CompDesc) * n_c);
    const size_t o_fp = pad(o_fd + sizeof(FillRegion) * std::max<size_t>(1, m.fills.size()));
    const size_t o_im = pad(o_fp + sizeof(FillRegion) * std::max<size_t>(1, m.fills_paged.size()));
    const size_t total = pad(o_im + 4 * std::max<size_t>(1, m.image.size()));
    int rc = ensure(ctx, ctx->modelblob, total);
    if (rc) return rc;
    std::vector<uint8_t> blob(total, 0);
    auto put = [&](size_t at, const void *p, size_t n) { if (n) std::memcpy(blob.data() + at, p, n); };
    put(o_hdr, m.header.data(), m.header.size());
    put(o_cd, m.comps.data(), sizeof(CompDesc) * m.comps.size());
    put(o_cp, m.comps_paged.data(), sizeof(CompDesc) * m.comps_paged.size());
    put(o_fd, m.fills.data(), sizeof(FillRegion) * m.fills.size());
    put(o_fp, m.fills_paged.data(), sizeof(FillRegion) * m.fills_paged.size());
    put(o_im, m.image.data(), 4 * m.image.size());
    // the blob of the previous launch may still be in use by queued kernels
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(ctx->modelblob.p, blob.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u8 *base = static_cast<u8 *>(ctx->modelblob.p);
    ModelDev md{};
    md.n = m.n, md.cend = m.cend, md.hbegin = m.hbegin, md.hend = m.hend, md.header_len = int(m.header.size());
    md.h_len = m.h_len, md.m_len = m.m_len;
    md.h_off = m.h_off, md.m_off = m.m_off, md.r_off = m.r_off, md.rt_off = m.rt_off;
    md.ws_bytes = m.ws_bytes;
    md.header = base + o_hdr;
    md.comps = reinterpret_cast<const CompDesc *>(base + o_cd);
    md.ctx_mode = m.ctx_mode, md.n_hash = m.n_hash;
    md.paged = 0, md.pool = nullptr, md.pool_pages = 0, md.pool_next = nullptr, md.pool_overflow = nullptr;
    out.dense = md;
    md.comps = reinterpret_cast<const CompDesc *>(base + o_cp);
    md.ws_bytes = m.ws_bytes_paged;
    md.paged = 1;
    out.paged = md;
    out.fills_dense = reinterpret_cast<const FillRegion *>(base + o_fd);
    out.fills_paged = reinterpret_cast<const FillRegion *>(base + o_fp);
    out.n_fills_dense = int(m.fills.size()), out.n_fills_paged = int(m.fills_paged.size());
    out.image = reinterpret_cast<const u32 *>(base + o_im);
    return ZPAQGPU_OK;
}

// Where the model tables of a batch live: dense slots (one wave = as many blocks as fit) or, for
// chain-shaped models whose dense tables do not fit, page tables plus a shared page pool.
struct TablePlan {
    bool paged = false;
    int slots = 0;        // blocks per wave
    u64 stride = 0;       // workspace bytes per block
    u64 pool_bytes = 0;
};

// round_cap: how many blocks the codec kernel holds at once (SMs x blocks per CTA; 0 = do not care).  A dense
// wave larger than that runs its CTAs in two rounds, the second nearly empty: a batch of several waves is
// therefore cut into as few rounds as its size needs, all of the same size (8 192 blocks at -m3: memory
// would take 1 127 per wave = 161 encoder CTAs on 148 SMs; 8 waves of 1 024 = 147 CTAs each instead).
int plan_tables(zpaqgpu_ctx *ctx, const Model &m, int n_blocks, size_t other_bytes, bool chain, bool force_dense,
                int round_cap, TablePlan &tp) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    free_b += ctx->workspace.cap + ctx->pool.cap;  // our own cached buffers can be reused
    u64 budget = ctx->ws_limit ? ctx->ws_limit : u64(double(free_b) * 0.80);
    if (!ctx->ws_limit && budget > other_bytes) budget -= std::min<u64>(other_bytes, budget / 2);
    const u64 dense_slots = m.ws_bytes ? budget / m.ws_bytes : u64(n_blocks);
    const bool can_page = chain && m.ws_bytes_paged > 0 && !force_dense && ctx->table_mode != ZPAQGPU_TABLES_DENSE;
    // Dense waves are kept while one wave still holds enough blocks to fill the device (four per SM): a
    // second wave costs nothing then, and incompressible blocks -- which touch every line and would run a
    // page pool dry -- need no retry.  Paging is for models whose dense tables leave most SMs idle (-m4/-m5).
    const u64 enough = std::min<u64>(u64(n_blocks), 4ull * u64(ctx->sm_count));
    const bool want_page = can_page && (ctx->table_mode == ZPAQGPU_TABLES_PAGED || dense_slots < enough);
    if (want_page) {
        u64 ht_bytes = 0;
        for (const CompDesc &c : m.comps)
            if (c.type == C_ICM || c.type == C_ISSE) ht_bytes += c.ht_len;
        const u64 max_blocks = (budget / 2) / m.ws_bytes_paged;  // at most half the budget for page tables
        if (max_blocks >= 1) {
            const u64 slots = std::min<u64>(max_blocks, u64(n_blocks));
            u64 pool = std::min<u64>(budget - slots * m.ws_bytes_paged, slots * ht_bytes);
            pool = pool / kPageBytes * kPageBytes;
            if (pool / kPageBytes > 0xFFFFFFF0ull) pool = 0xFFFFFFF0ull * kPageBytes;
            if (pool >= kPageBytes * 16) {
                tp.paged = true, tp.slots = int(slots), tp.stride = m.ws_bytes_paged, tp.pool_bytes = pool;
                return ZPAQGPU_OK;
            }
        }
    }
    if (dense_slots < 1) {
        ctx->err = "model tables (" + std::to_string(m.ws_bytes) + " bytes per block) exceed the workspace budget";
        return ZPAQGPU_E_NOMEM;
    }
    const u64 slots = wave_slots(u64(n_blocks), dense_slots, u64(std::max(0, round_cap)));
    tp.paged = false, tp.slots = int(slots), tp.stride = m.ws_bytes, tp.pool_bytes = 0;
    return ZPAQGPU_OK;
}

// Clear and initialise the tables of one wave; returns the ModelDev to launch with.
int prepare_wave(zpaqgpu_ctx *ctx, const ModelOnDev &mod, const TablePlan &tp, int n, ModelDev &md) {
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(ctx->workspace.p, 0, u64(n) * tp.stride, st));
    FillArgs fa{static_cast<u8 *>(ctx->workspace.p), tp.stride, n,
                tp.paged ? mod.fills_paged : mod.fills_dense, tp.paged ? mod.n_fills_paged : mod.n_fills_dense,
                mod.image};
    launch_fill(fa, st);
    md = tp.paged ? mod.paged : mod.dense;
    if (tp.paged) {
        CK(cudaMemsetAsync(ctx->pool.p, 0, tp.pool_bytes, st));
        CK(cudaMemsetAsync(static_cast<u8 *>(ctx->misc.p) + 16, 0, 8, st));
        md.pool = static_cast<u8 *>(ctx->pool.p);
        md.pool_pages = u32(tp.pool_bytes / kPageBytes);
        md.pool_next = reinterpret_cast<u32 *>(static_cast<u8 *>(ctx->misc.p) + 16);
        md.pool_overflow = md.pool_next + 1;
    }
    return ZPAQGPU_OK;
}

// After a paged wave: pages used, and whether the pool ran dry.
int wave_pool_status(zpaqgpu_ctx *ctx, u32 &used, bool &overflow) {
    u32 two[2] = {0, 0};
    CK(cudaMemcpyAsync(two, static_cast<u8 *>(ctx->misc.p) + 16, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    used = two[0], overflow = two[1] != 0;
    return ZPAQGPU_OK;
}

bool use_chain(const zpaqgpu_ctx *ctx, const Model &m) {
    if (ctx->kernel_pref == ZPAQGPU_KERNEL_GENERIC) return false;
    return m.is_chain;
}

// Blocks per CTA of a launch of n blocks when a CTA can hold at most `most` (one CTA per SM: shared memory):
// as many as spread the launch over all SMs; when that is more than a CTA holds, the CTAs run in rounds, and
// the rounds are made equal instead of a full one followed by a nearly empty one (-m5 encoder: 1 024 blocks,
// 6 per CTA at most = 171 CTAs on 148 SMs; 4 per CTA = two rounds of 128).
int per_cta(const zpaqgpu_ctx *ctx, int n, int most) { return blocks_per_cta(n, most, ctx->sm_count); }

int pick_warps_per_cta(const zpaqgpu_ctx *ctx, const Model &m, int n_resident) {
    return per_cta(ctx, n_resident, chain_max_warps_per_cta(m));
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// ------------------------------------------------------------------------------------------
// Compression job: blocks of segments, plaintext already on the device.
// ------------------------------------------------------------------------------------------
int run_compress(zpaqgpu_ctx *ctx, CompressJob &job) {
    Range nvtx_job("zpaqgpu:compress_job");
    const Model &m = *job.model;
    const int n_blocks = int(job.blocks.size()), n_segs = int(job.segs.size());
    ctx->stats = zpaqgpu_stats{};
    ctx->stats.workspace_bytes_per_block = m.ws_bytes;
    const bool store = m.n == 0;
    const bool chain = !store && use_chain(ctx, m);
    if (!store && ctx->kernel_pref == ZPAQGPU_KERNEL_CHAIN && !m.is_chain) {
        ctx->err = "model does not have the ICM/ISSE-chain shape";
        return ZPAQGPU_E_UNSUPPORTED;
    }
    ctx->stats.kernel = store ? 0 : (chain ? ZPAQGPU_KERNEL_CHAIN : ZPAQGPU_KERNEL_GENERIC);
    cudaStream_t st = ctx->stream;

    // framing bytes in front of each payload (compressor.v:150-181, :217-235)
    std::vector<uint8_t> pre;
    std::vector<PackSeg> pack(static_cast<size_t>(n_segs));
    std::vector<EncSeg> esegs(static_cast<size_t>(n_segs));
    std::vector<ShaJob> sha(static_cast<size_t>(n_segs));
    std::vector<u64> caps(static_cast<size_t>(n_segs), 0);
    for (int attempt = 0; attempt < 3; ++attempt) {
        pre.clear();
        u64 arena_bytes = 0;
        for (int b = 0; b < n_blocks; ++b) {
            const EncBlock &blk = job.blocks[size_t(b)];
            for (u32 k = 0; k < blk.n_seg; ++k) {
                const u32 s = blk.first_seg + k;
                const SegSpec &sp = job.segs[s];
                PackSeg &ps = pack[s];
                ps.pre_off = pre.size();
                if (k == 0) pre.insert(pre.end(), m.block_prefix.begin(), m.block_prefix.end());
                pre.push_back(1);
                for (const char *c = sp.name ? sp.name : ""; *c; ++c) pre.push_back(uint8_t(*c));
                pre.push_back(0);
                for (const char *c = sp.comment ? sp.comment : ""; *c; ++c) pre.push_back(uint8_t(*c));
                pre.push_back(0);
                pre.push_back(0);
                ps.pre_len = u32(pre.size() - ps.pre_off);
                ps.store = store ? 1 : 0;
                ps.in_off = sp.in_off, ps.in_len = sp.in_len;
                ps.flags = sp.called ? 1u : 0u;
                ps.last = (k + 1 == blk.n_seg) ? 1 : 0;
                // The reference's squash is inverted beyond |d| >= 1018: a model that has
                // become very sure of itself can expand a block by 1.5x.  Twice the input covers that;
                // anything larger is caught by the sizing retry below.
                if (attempt == 0) caps[s] = store ? 0 : 2 * sp.in_len + 1024;
                ps.pay_off = arena_bytes, ps.pay_cap = caps[s];
                arena_bytes += align_up(caps[s], 256);
                EncSeg &es = esegs[s];
                es.in_off = sp.in_off, es.in_len = sp.in_len, es.pay_off = ps.pay_off, es.pay_cap = caps[s];
                es.flags = ps.flags, es.pad = 0;
                sha[s].off = sp.in_off, sha[s].len = sp.in_len;
            }
        }
        // descriptor blob
        const size_t o_blocks = 0;
        const size_t o_esegs = align_up(o_blocks + sizeof(EncBlock) * size_t(n_blocks), 16);
        const size_t o_pack = align_up(o_esegs + sizeof(EncSeg) * size_t(n_segs), 16);
        const size_t o_sha = align_up(o_pack + sizeof(PackSeg) * size_t(n_segs), 16);
        const size_t o_order = align_up(o_sha + sizeof(ShaJob) * size_t(n_segs), 16);
        const size_t o_pre = align_up(o_order + sizeof(u32) * size_t(n_blocks), 16);
        const size_t desc_bytes = align_up(o_pre + pre.size() + 16, 16);
        int rc;
        if ((rc = ensure(ctx, ctx->desc, desc_bytes))) return rc;
        if ((rc = ensure(ctx, ctx->arena, std::max<u64>(arena_bytes, 256)))) return rc;
        if ((rc = ensure(ctx, ctx->pay_len, 8 * size_t(n_segs) + 16))) return rc;
        if ((rc = ensure(ctx, ctx->digests, 20 * size_t(n_segs) + 16))) return rc;
        if ((rc = ensure(ctx, ctx->seg_size, 8 * size_t(n_segs) + 16))) return rc;
        std::vector<uint8_t> blob(desc_bytes, 0);
        std::memcpy(blob.data() + o_blocks, job.blocks.data(), sizeof(EncBlock) * size_t(n_blocks));
        std::memcpy(blob.data() + o_esegs, esegs.data(), sizeof(EncSeg) * size_t(n_segs));
        std::memcpy(blob.data() + o_pack, pack.data(), sizeof(PackSeg) * size_t(n_segs));
        std::memcpy(blob.data() + o_sha, sha.data(), sizeof(ShaJob) * size_t(n_segs));
        {   // dispatch order: largest blocks first (identity when all blocks have one size)
            std::vector<u32> order(static_cast<size_t>(n_blocks));
            std::vector<u64> weight(static_cast<size_t>(n_blocks), 0);
            for (int b = 0; b < n_blocks; ++b) {
                order[size_t(b)] = u32(b);
                const EncBlock &blk = job.blocks[size_t(b)];
                for (u32 k = 0; k < blk.n_seg; ++k) weight[size_t(b)] += job.segs[blk.first_seg + k].in_len;
            }
            std::stable_sort(order.begin(), order.end(), [&](u32 x, u32 y) { return weight[x] > weight[y]; });
            std::memcpy(blob.data() + o_order, order.data(), sizeof(u32) * size_t(n_blocks));
        }
        if (!pre.empty()) std::memcpy(blob.data() + o_pre, pre.data(), pre.size());
        CK(cudaMemcpyAsync(ctx->desc.p, blob.data(), desc_bytes, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // blob is a stack-lifetime host buffer
        u8 *dbase = static_cast<u8 *>(ctx->desc.p);
        const EncBlock *d_blocks = reinterpret_cast<const EncBlock *>(dbase + o_blocks);
        const EncSeg *d_esegs = reinterpret_cast<const EncSeg *>(dbase + o_esegs);
        const PackSeg *d_pack = reinterpret_cast<const PackSeg *>(dbase + o_pack);
        const ShaJob *d_sha = reinterpret_cast<const ShaJob *>(dbase + o_sha);
        const u32 *d_order = reinterpret_cast<const u32 *>(dbase + o_order);
        const u8 *d_pre = dbase + o_pre;

        // SHA-1 of the plaintext runs beside the codec on the side stream (compressor.v:284)
        CK(cudaEventRecord(ctx->ev_main, st));
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_main, 0));
        CK(cudaEventRecord(ctx->ev[4], ctx->side_stream));
        launch_sha1(job.d_in, d_sha, n_segs, static_cast<u8 *>(ctx->digests.p), ctx->side_stream);
        CK(cudaEventRecord(ctx->ev[5], ctx->side_stream));
        CK(cudaEventRecord(ctx->ev_side, ctx->side_stream));
        ctx->stats.launches += 1;

        float init_ms = 0, codec_ms = 0;
        if (!store) {
            ModelOnDev mod;
            if ((rc = upload_model(ctx, m, mod))) return rc;
            if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
            bool force_dense = false;
            for (int table_try = 0; table_try < 2; ++table_try) {
                TablePlan tp;
                const int round_cap = chain ? ctx->sm_count * encpipe_max_blocks_per_cta(m) : 0;
                if ((rc = plan_tables(ctx, m, n_blocks, arena_bytes, chain, force_dense, round_cap, tp))) return rc;
                if ((rc = ensure_tables(ctx, u64(tp.slots) * tp.stride, tp.paged ? tp.pool_bytes : 0))) return rc;
                // encoder: three warps per block, as many blocks per CTA as spreads the wave over all SMs
                int wpc = 4;
                if (chain) wpc = per_cta(ctx, tp.slots, encpipe_max_blocks_per_cta(m));
                ctx->stats.warps_per_cta = chain ? wpc * 3 : wpc;
                ctx->stats.paged = tp.paged ? 1 : 0;
                ctx->stats.workspace_bytes_per_block = tp.stride;
                bool overflow = false;
                for (int first = 0; first < n_blocks && !overflow; first += tp.slots) {
                    const int n = std::min(tp.slots, n_blocks - first);
                    Range nvtx_wave("zpaqgpu:encode_wave");
                    CK(cudaEventRecord(ctx->ev[0], st));
                    ModelDev md;
                    if ((rc = prepare_wave(ctx, mod, tp, n, md))) return rc;
                    CK(cudaEventRecord(ctx->ev[1], st));
                    EncodeArgs ea;
                    ea.model = md, ea.tables = ctx->tables;
                    ea.workspace = static_cast<u8 *>(ctx->workspace.p);
                    ea.in = job.d_in, ea.arena = static_cast<u8 *>(ctx->arena.p);
                    ea.blocks = d_blocks, ea.segs = d_esegs, ea.pay_len = static_cast<u64 *>(ctx->pay_len.p);
                    ea.order = d_order, ea.first_block = first, ea.n_blocks = n;
                    ea.flags = (ctx->enc_flags & 1) | (ctx->ahead ? 0 : 4);
                    if (chain) {
                        if (!launch_encode_pipe3(m, ea, wpc, st)) {
                            ctx->err = "no chain kernel instantiation for this model";
                            return ZPAQGPU_E_UNSUPPORTED;
                        }
                    } else if (ctx->generic_warp && genwarp_supports(m)) {
                        if (!launch_encode_genwarp(ea, st)) return ctx->err = "generic warp kernel launch", ZPAQGPU_E_CUDA;
                    } else {
                        k_encode_generic<<<(n + 3) / 4, 128, 0, st>>>(ea);
                    }
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(ctx->ev[2], st));
                    CK(cudaEventSynchronize(ctx->ev[2]));
                    init_ms += elapsed(ctx->ev[0], ctx->ev[1]);
                    codec_ms += elapsed(ctx->ev[1], ctx->ev[2]);
                    ctx->stats.launches += 2 + (m.fills.empty() ? 0 : 1) + (tp.paged ? 2 : 0);
                    ctx->stats.codec_launches += 1;
                    ctx->stats.waves += 1;
                    if (tp.paged) {
                        u32 used = 0;
                        if ((rc = wave_pool_status(ctx, used, overflow))) return rc;
                        ctx->stats.pool_bytes_used = std::max<u64>(ctx->stats.pool_bytes_used, u64(used) * kPageBytes);
                    }
                }
                if (!overflow) break;
                // the page pool ran dry (incompressible data touches every line): dense waves instead
                force_dense = true;
                ctx->stats.retries += 1;
                if (table_try == 1) {
                    ctx->err = "page pool overflow with dense tables";
                    return ZPAQGPU_E_CUDA;
                }
            }
        }
        ctx->stats.init_ms += init_ms, ctx->stats.codec_ms += codec_ms;
        // assemble: [prefix][payload][00000000 FD sha1][FF] per segment (compressor.v:380-395, :409)
        CK(cudaStreamWaitEvent(st, ctx->ev_side, 0));
        CK(cudaEventRecord(ctx->ev[2], st));
        PackArgs pa;
        pa.segs = d_pack, pa.blocks = d_blocks, pa.n_blocks = n_blocks, pa.pre = d_pre, pa.in = job.d_in;
        pa.arena = static_cast<const u8 *>(ctx->arena.p), pa.pay_len = static_cast<const u64 *>(ctx->pay_len.p);
        pa.digests = static_cast<const u8 *>(ctx->digests.p), pa.seg_size = static_cast<u64 *>(ctx->seg_size.p);
        pa.out_off = job.d_out_off, pa.out = job.d_out, pa.out_cap = job.out_cap;
        launch_pack(pa, n_segs, st);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[3], st));
        ctx->stats.launches += 3;
        // read back the total and the payload sizes to detect slot overflow
        if ((rc = ensure_pinned(ctx, 8 * size_t(n_segs) + 64))) return rc;
        u64 *h_total = static_cast<u64 *>(ctx->pinned);
        u64 *h_pay = h_total + 1;
        CK(cudaMemcpyAsync(h_total, job.d_out_off + n_blocks, 8, cudaMemcpyDeviceToHost, st));
        if (!store) CK(cudaMemcpyAsync(h_pay, ctx->pay_len.p, 8 * size_t(n_segs), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats.pack_ms += elapsed(ctx->ev[2], ctx->ev[3]);
        ctx->stats.sha1_ms += elapsed(ctx->ev[4], ctx->ev[5]);
        bool overflow = false;
        if (!store)
            for (int s = 0; s < n_segs; ++s)
                if (h_pay[s] > caps[size_t(s)]) {
                    if (!overflow)
                        ctx->err = "payload of segment " + std::to_string(s) + " needs " + std::to_string(h_pay[s]) +
                                   " bytes, slot had " + std::to_string(caps[size_t(s)]) + " (retried)";
                    caps[size_t(s)] = h_pay[s] + 64, overflow = true;
                }
        if (!overflow) {
            job.total = *h_total;
            job.fits = job.total <= job.out_cap;
            return ZPAQGPU_OK;
        }
        ctx->stats.retries += 1;
    }
    ctx->err = "payload slot sizing did not converge";
    return ZPAQGPU_E_CUDA;
}

std::string size_comment(u64 n) { return std::to_string(n) + " bytes"; }

// Host buffers -> archive bytes in ctx->out and block offsets in ctx->out_off (device): the upload and the
// compression job, without the copy back.  *total receives the archive size.
int compress_stage(zpaqgpu_ctx *ctx, const Model &m, const uint8_t *in, const uint64_t *in_off, int n_blocks,
                   const char *const *names, const char *const *comments, u64 *total) {
    Range nvtx_call("zpaqgpu:compress_stage");
    CK(cudaSetDevice(ctx->device));
    const u64 base = in_off[0], total_in = in_off[n_blocks] - base;
    if (total_in > 0 && !in) return ZPAQGPU_E_ARG;
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->in, std::max<u64>(total_in, 16)))) return rc;
    CK(cudaEventRecord(ctx->ev[6], st));
    if (total_in) CK(cudaMemcpyAsync(ctx->in.p, in + base, total_in, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[7], st));
    CompressJob job;
    job.model = &m;
    job.blocks.resize(size_t(n_blocks));
    job.segs.resize(size_t(n_blocks));
    u64 worst = 0;
    for (int b = 0; b < n_blocks; ++b) {
        if (in_off[b + 1] < in_off[b]) return ZPAQGPU_E_ARG;
        job.blocks[size_t(b)] = EncBlock{u32(b), 1};
        SegSpec &sp = job.segs[size_t(b)];
        sp.name = names ? names[b] : nullptr;
        sp.comment = comments ? comments[b] : nullptr;
        sp.in_off = in_off[b] - base, sp.in_len = in_off[b + 1] - in_off[b];
        sp.called = true;  // cmd/main.v:305 always calls compress() at least once
        worst += sp.in_len + sp.in_len / 2 + 2048 + std::strlen(sp.name ? sp.name : "") +
                 std::strlen(sp.comment ? sp.comment : "");
    }
    // the assembled archive is produced on the device and copied out in one piece
    const u64 dcap = worst;
    if ((rc = ensure(ctx, ctx->out, dcap))) return rc;
    if ((rc = ensure(ctx, ctx->out_off, 8 * size_t(n_blocks + 1)))) return rc;
    job.d_in = static_cast<const u8 *>(ctx->in.p);
    job.d_out = static_cast<u8 *>(ctx->out.p), job.out_cap = ctx->out.cap;
    job.d_out_off = static_cast<u64 *>(ctx->out_off.p);
    if ((rc = run_compress(ctx, job))) return rc;
    ctx->stats.h2d_ms = elapsed(ctx->ev[6], ctx->ev[7]);
    if (!job.fits) {
        // the device buffer was sized from a worst-case guess; grow it and assemble again
        if ((rc = ensure(ctx, ctx->out, job.total))) return rc;
        job.d_out = static_cast<u8 *>(ctx->out.p), job.out_cap = ctx->out.cap;
        if ((rc = run_compress(ctx, job))) return rc;
    }
    *total = job.total;
    return ZPAQGPU_OK;
}

// The staged archive back to the host: `total` bytes to out, the n_blocks+1 offsets (plus `shift`) to out_off.
int compress_fetch(zpaqgpu_ctx *ctx, int n_blocks, u64 total, uint8_t *out, uint64_t *out_off, u64 shift) {
    Range nvtx_call("zpaqgpu:compress_fetch");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(out_off, ctx->out_off.p, 8 * size_t(n_blocks + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[6], st));
    if (total) CK(cudaMemcpyAsync(out, ctx->out.p, total, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[7], st));
    CK(cudaStreamSynchronize(st));
    ctx->stats.d2h_ms = elapsed(ctx->ev[6], ctx->ev[7]);
    if (shift)
        for (int b = 0; b <= n_blocks; ++b) out_off[b] += shift;
    return ZPAQGPU_OK;
}

int compress_host(zpaqgpu_ctx *ctx, const Model &m, const uint8_t *in, const uint64_t *in_off, int n_blocks,
                  const char *const *names, const char *const *comments, uint8_t *out, uint64_t out_cap,
                  uint64_t *out_off, uint64_t *out_need) {
    if (!ctx || n_blocks < 0 || (n_blocks > 0 && (!in_off || !out_off))) return ZPAQGPU_E_ARG;
    if (n_blocks == 0) {
        if (out_off) out_off[0] = 0;
        if (out_need) *out_need = 0;
        return ZPAQGPU_OK;
    }
    u64 total = 0;
    int rc = compress_stage(ctx, m, in, in_off, n_blocks, names, comments, &total);
    if (rc) return rc;
    if (out_need) *out_need = total;
    if (total > out_cap) {
        CK(cudaMemcpyAsync(out_off, ctx->out_off.p, 8 * size_t(n_blocks + 1), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return ZPAQGPU_E_NOSPACE;
    }
    if (total && !out) return ZPAQGPU_E_ARG;
    return compress_fetch(ctx, n_blocks, total, out, out_off, 0);
}

// ------------------------------------------------------------------------------------------
// Decompression job over an archive resident on the device.
// ------------------------------------------------------------------------------------------
struct DecCandidate {
    u64 start;       // offset of the level byte (just after the locator)
    u64 payload;     // offset of the first segment marker
    int group = -1;  // index into models, -1: header rejected
    u64 hint = 0;    // plaintext capacity to reserve
    bool hinted = false;  // the capacity comes from the segment comment
    u64 out_off = 0;
    DecBlockOut res{};
};

struct DecompressJob {
    const u8 *d_arc;
    u64 arc_len;
    std::vector<DecCandidate> cand;
    std::vector<Model> models;
    u8 *d_plain;     // plaintext arena on the device
    std::vector<DecSegRec> recs;   // sorted by (block, index)
    std::vector<i32> sha_ok;       // parallel to recs
};

int run_decompress(zpaqgpu_ctx *ctx, DecompressJob &job, bool caller_owns_plain) {
    Range nvtx_job("zpaqgpu:decompress_job");
    cudaStream_t st = ctx->stream;
    const int n = int(job.cand.size());
    ctx->stats = zpaqgpu_stats{};
    int rc;
    bool dense_only = false;       // set after a paged attempt ran out of pool pages
    bool pool_overflowed = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
        // plaintext slots, back to back in candidate order
        u64 plain_bytes = 0;
        std::vector<DecBlock> blocks(static_cast<size_t>(n));
        for (int i = 0; i < n; ++i) {
            DecCandidate &c = job.cand[size_t(i)];
            if (!caller_owns_plain) c.out_off = plain_bytes;
            blocks[size_t(i)] = DecBlock{c.payload, c.out_off, c.hint};
            plain_bytes += c.hint;
        }
        if (!caller_owns_plain) {
            if ((rc = ensure(ctx, ctx->plain, std::max<u64>(plain_bytes, 16)))) return rc;
            job.d_plain = static_cast<u8 *>(ctx->plain.p);
        }
        u32 seg_cap = u32(std::max(64, n * 2 + 64));
        for (int seg_try = 0; seg_try < 2; ++seg_try) {
            const size_t o_dorder = align_up(sizeof(DecBlock) * size_t(n), 16);
            if ((rc = ensure(ctx, ctx->desc, o_dorder + sizeof(u32) * size_t(n) + 16))) return rc;
            if ((rc = ensure(ctx, ctx->results, sizeof(DecBlockOut) * size_t(n) + 16))) return rc;
            if ((rc = ensure(ctx, ctx->seg_recs, sizeof(DecSegRec) * size_t(seg_cap) + 16))) return rc;
            if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
            CK(cudaMemcpyAsync(ctx->desc.p, blocks.data(), sizeof(DecBlock) * size_t(n), cudaMemcpyHostToDevice, st));
            {   // dispatch order: inside every run of blocks with one model, largest plaintext first
                std::vector<u32> order(static_cast<size_t>(n));
                for (int k = 0; k < n; ++k) order[size_t(k)] = u32(k);
                int a = 0;
                while (a < n) {
                    int b = a;
                    while (b < n && job.cand[size_t(b)].group == job.cand[size_t(a)].group) ++b;
                    std::stable_sort(order.begin() + a, order.begin() + b,
                                     [&](u32 x, u32 y) { return blocks[x].out_cap > blocks[y].out_cap; });
                    a = b;
                }
                CK(cudaMemcpyAsync(static_cast<u8 *>(ctx->desc.p) + o_dorder, order.data(), sizeof(u32) * size_t(n),
                                   cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));  // order is a stack-lifetime host buffer
            }
            CK(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
            CK(cudaMemsetAsync(ctx->results.p, 0, sizeof(DecBlockOut) * size_t(n), st));
            CK(cudaStreamSynchronize(st));
            // one launch group per distinct model; candidates of a group are contiguous runs
            for (size_t g = 0; g < job.models.size(); ++g) {
                const Model &m = job.models[g];
                ctx->stats.workspace_bytes_per_block = m.ws_bytes;
                ModelOnDev mod;
                if ((rc = upload_model(ctx, m, mod))) return rc;
                const bool store = m.n == 0;
                const bool chain = !store && use_chain(ctx, m);
                if (!store && ctx->kernel_pref == ZPAQGPU_KERNEL_CHAIN && !m.is_chain) {
                    ctx->err = "model does not have the ICM/ISSE-chain shape";
                    return ZPAQGPU_E_UNSUPPORTED;
                }
                if (!store) ctx->stats.kernel = chain ? ZPAQGPU_KERNEL_CHAIN : ZPAQGPU_KERNEL_GENERIC;
                int i = 0;
                while (i < n) {
                    if (job.cand[size_t(i)].group != int(g)) { ++i; continue; }
                    int j = i;
                    while (j < n && job.cand[size_t(j)].group == int(g)) ++j;
                    const int run = j - i;
                    for (int once = 0; once < 1; ++once) {
                        TablePlan tp;
                        tp.slots = run;
                        if (!store) {
                            const int round_cap =
                                !chain ? 0
                                       : ctx->sm_count * ((ctx->decoder == 2 && tree2_supports(m)) ? tree2_max_pairs_per_cta(m)
                                                                                                    : chain_max_warps_per_cta(m));
                            if ((rc = plan_tables(ctx, m, run, plain_bytes, chain, dense_only, round_cap, tp))) return rc;
                            if ((rc = ensure_tables(ctx, u64(tp.slots) * tp.stride, tp.paged ? tp.pool_bytes : 0)))
                                return rc;
                            ctx->stats.paged = tp.paged ? 1 : 0;
                            ctx->stats.workspace_bytes_per_block = tp.stride;
                        }
                        const bool tree2 = chain && ctx->decoder == 2 && tree2_supports(m);
                        int wpc = chain ? pick_warps_per_cta(ctx, m, tp.slots) : 4;
                        if (tree2) wpc = per_cta(ctx, tp.slots, tree2_max_pairs_per_cta(m));
                        ctx->stats.warps_per_cta = tree2 ? 2 * wpc : wpc;
                        bool overflow = false;
                        for (int first = i; first < j && !overflow; first += tp.slots) {
                            const int cnt = std::min(tp.slots, j - first);
                            Range nvtx_wave("zpaqgpu:decode_wave");
                            DecodeArgs da;
                            da.tables = ctx->tables;
                            da.workspace = static_cast<u8 *>(ctx->workspace.p);
                            da.arc = job.d_arc, da.arc_len = job.arc_len, da.out = job.d_plain;
                            da.blocks = static_cast<const DecBlock *>(ctx->desc.p);
                            da.results = static_cast<DecBlockOut *>(ctx->results.p);
                            da.seg_recs = static_cast<DecSegRec *>(ctx->seg_recs.p);
                            da.seg_count = static_cast<u32 *>(ctx->misc.p);
                            da.seg_cap = seg_cap, da.first_block = first, da.n_blocks = cnt;
                            da.order = reinterpret_cast<const u32 *>(static_cast<const u8 *>(ctx->desc.p) + o_dorder);
                            da.flags = (ctx->spec_probe ? 1 : 0) | (ctx->guess << 8) | (ctx->pull_how << 12);
                            CK(cudaEventRecord(ctx->ev[0], st));
                            if (store) {
                                da.model = mod.dense;
                                CK(cudaEventRecord(ctx->ev[1], st));
                                launch_decode_store(da, st);
                            } else {
                                if ((rc = prepare_wave(ctx, mod, tp, cnt, da.model))) return rc;
                                CK(cudaEventRecord(ctx->ev[1], st));
                                if (chain) {
                                    const bool ok = tree2 ? launch_decode_tree2(m, da, wpc, st)
                                                          : launch_decode_chain(m, da, wpc, ctx->decoder != 0, st);
                                    if (!ok) {
                                        ctx->err = "no chain kernel instantiation for this model";
                                        return ZPAQGPU_E_UNSUPPORTED;
                                    }
                                } else if (ctx->generic_warp && genwarp_supports(m)) {
                                    if (!launch_decode_genwarp(da, st))
                                        return ctx->err = "generic warp kernel launch", ZPAQGPU_E_CUDA;
                                } else {
                                    k_decode_generic<<<(cnt + 3) / 4, 128, 0, st>>>(da);
                                }
                                ctx->stats.launches += 1 + (m.fills.empty() ? 0 : 1) + (tp.paged ? 2 : 0);
                            }
                            CK(cudaGetLastError());
                            CK(cudaEventRecord(ctx->ev[2], st));
                            CK(cudaEventSynchronize(ctx->ev[2]));
                            ctx->stats.init_ms += elapsed(ctx->ev[0], ctx->ev[1]);
                            ctx->stats.codec_ms += elapsed(ctx->ev[1], ctx->ev[2]);
                            ctx->stats.launches += 1;
                            ctx->stats.codec_launches += 1;
                            ctx->stats.waves += 1;
                            if (!store && tp.paged) {
                                u32 used = 0;
                                if ((rc = wave_pool_status(ctx, used, overflow))) return rc;
                                ctx->stats.pool_bytes_used =
                                    std::max<u64>(ctx->stats.pool_bytes_used, u64(used) * kPageBytes);
                            }
                        }
                        if (overflow) pool_overflowed = true;
                        break;
                    }
                    if (pool_overflowed) break;
                    i = j;
                }
                if (pool_overflowed) break;
            }
            if (pool_overflowed) break;
            // results
            if ((rc = ensure_pinned(ctx, sizeof(DecBlockOut) * size_t(n) + 64))) return rc;
            u32 *h_count = static_cast<u32 *>(ctx->pinned);
            DecBlockOut *h_res = reinterpret_cast<DecBlockOut *>(static_cast<u8 *>(ctx->pinned) + 64);
            CK(cudaMemcpyAsync(h_count, ctx->misc.p, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h_res, ctx->results.p, sizeof(DecBlockOut) * size_t(n), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const u32 count = *h_count;
            for (int i = 0; i < n; ++i) job.cand[size_t(i)].res = h_res[i];
            if (count > seg_cap) {  // more segments than record slots: run again with room for all
                seg_cap = count + 64;
                ctx->stats.retries += 1;
                continue;
            }
            job.recs.resize(count);
            if (count) {
                CK(cudaMemcpy(job.recs.data(), ctx->seg_recs.p, sizeof(DecSegRec) * size_t(count), cudaMemcpyDeviceToHost));
            }
            break;
        }
        if (pool_overflowed) {
            // incompressible data touched more lines than the page pool holds: the whole call is
            // repeated with dense tables in as many waves as needed
            if (dense_only) {
                ctx->err = "page pool overflow with dense tables";
                return ZPAQGPU_E_CUDA;
            }
            dense_only = true, pool_overflowed = false;
            ctx->stats.retries += 1;
            --attempt;
            continue;
        }
        bool overflow = false;
        for (int i = 0; i < n; ++i) {
            DecCandidate &c = job.cand[size_t(i)];
            if (c.group >= 0 && c.res.out_len > c.hint) {
                if (caller_owns_plain) {
                    ctx->err = "block " + std::to_string(i) + " decodes to " + std::to_string(c.res.out_len) +
                               " bytes, more than its output slot";
                    return ZPAQGPU_E_NOSPACE;
                }
                overflow = true;
            }
            if (c.group >= 0) c.hint = std::max<u64>(c.res.out_len, 1);
        }
        if (!overflow) break;
        if (attempt == 1) {
            ctx->err = "plaintext slot sizing did not converge";
            return ZPAQGPU_E_CUDA;
        }
        ctx->stats.retries += 1;
    }
    std::sort(job.recs.begin(), job.recs.end(), [](const DecSegRec &a, const DecSegRec &b) {
        return a.block != b.block ? a.block < b.block : a.index < b.index;
    });
    // SHA-1 of every segment's plaintext against the stored digest (decompressor.v:608-628)
    const int ns = int(job.recs.size());
    job.sha_ok.assign(size_t(ns), -1);
    if (ns) {
        std::vector<ShaJob> jobs(static_cast<size_t>(ns));
        for (int k = 0; k < ns; ++k) jobs[size_t(k)] = ShaJob{job.recs[size_t(k)].out_off, job.recs[size_t(k)].out_len};
        if ((rc = ensure(ctx, ctx->seg_size, sizeof(ShaJob) * size_t(ns) + 16))) return rc;
        if ((rc = ensure(ctx, ctx->digests, 24 * size_t(ns) + 16))) return rc;
        if ((rc = ensure(ctx, ctx->seg_recs, sizeof(DecSegRec) * size_t(ns) + 16))) return rc;
        CK(cudaMemcpyAsync(ctx->seg_size.p, jobs.data(), sizeof(ShaJob) * size_t(ns), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->seg_recs.p, job.recs.data(), sizeof(DecSegRec) * size_t(ns), cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ctx->ev[4], st));
        launch_sha1(job.d_plain, static_cast<const ShaJob *>(ctx->seg_size.p), ns, static_cast<u8 *>(ctx->digests.p), st);
        i32 *d_ok = reinterpret_cast<i32 *>(static_cast<u8 *>(ctx->digests.p) + 20 * size_t(ns));
        d_ok = reinterpret_cast<i32 *>(align_up(reinterpret_cast<uintptr_t>(d_ok), 4));
        launch_sha_compare(job.d_arc, job.arc_len, static_cast<const DecSegRec *>(ctx->seg_recs.p), ns,
                           static_cast<const u8 *>(ctx->digests.p), d_ok, st);
        CK(cudaEventRecord(ctx->ev[5], st));
        CK(cudaMemcpyAsync(job.sha_ok.data(), d_ok, 4 * size_t(ns), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats.sha1_ms = elapsed(ctx->ev[4], ctx->ev[5]);
        ctx->stats.launches += 2;
    }
    return ZPAQGPU_OK;
}

// "<N> bytes" comment convention of cmd/main.v:295-303 as a capacity hint
u64 comment_hint(const uint8_t *arc, u64 len, u64 payload) {
    u64 p = payload;
    if (p >= len || arc[p] != 1) return 0;
    ++p;
    while (p < len && arc[p]) ++p;  // filename
    ++p;
    u64 v = 0;
    int digits = 0;
    while (p < len && arc[p] >= '0' && arc[p] <= '9' && digits < 15) v = v * 10 + (arc[p] - '0'), ++p, ++digits;
    if (!digits || p + 6 >= len || std::memcmp(arc + p, " bytes", 7) != 0) return 0;
    return v;
}

// Steps 1-4 of decoding a whole archive: the archive goes to the device, every block is found and
// decoded into the plaintext arena (ctx->plain), and the segments are listed the way repeated
// find_block / find_filename calls would meet them.  *status carries what the walk ended with.
int decode_archive_dev(zpaqgpu_ctx *ctx, const uint8_t *arc, u64 len, std::vector<DecodedSeg> &segs,
                       const u8 **d_plain, int *status_out, u64 *total_out) {
    segs.clear();
    Range nvtx_call("zpaqgpu:decode_archive");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    // 1. archive to the device, locator scan
    if ((rc = ensure(ctx, ctx->in, len))) return rc;
    CK(cudaEventRecord(ctx->ev[6], st));
    CK(cudaMemcpyAsync(ctx->in.p, arc, len, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[7], st));
    u32 dcap = 1u << 16;
    std::vector<u64> starts;
    for (;;) {
        if ((rc = ensure(ctx, ctx->results, 8 * size_t(dcap)))) return rc;
        if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
        CK(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
        launch_find_blocks(static_cast<const u8 *>(ctx->in.p), len, static_cast<u64 *>(ctx->results.p), dcap,
                           static_cast<u32 *>(ctx->misc.p), st);
        CK(cudaGetLastError());
        u32 count = 0;
        CK(cudaMemcpyAsync(&count, ctx->misc.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (count > dcap) { dcap = count + 1024; continue; }
        starts.resize(count);
        if (count) CK(cudaMemcpy(starts.data(), ctx->results.p, 8 * size_t(count), cudaMemcpyDeviceToHost));
        break;
    }
    const float h2d_ms = elapsed(ctx->ev[6], ctx->ev[7]);
    std::sort(starts.begin(), starts.end());
    // 2. headers (decompressor.v:257-342).  Every candidate is decoded optimistically; candidates
    //    that turn out to lie inside an earlier block are dropped in step 4.
    DecompressJob job;
    job.d_arc = static_cast<const u8 *>(ctx->in.p), job.arc_len = len;
    std::map<std::vector<uint8_t>, int> group_of;
    for (u64 s : starts) {
        DecCandidate c;
        c.start = s;
        Model m;
        u64 used = 0;
        const int hrc = model_from_archive(arc + s, len - s, m, &used);
        c.payload = s + used;
        if (hrc == ZPAQGPU_OK) {
            auto it = group_of.find(m.header);
            if (it == group_of.end()) {
                it = group_of.emplace(m.header, int(job.models.size())).first;
                job.models.push_back(m);
            }
            c.group = it->second;
            const u64 hint = comment_hint(arc, len, c.payload);
            c.hint = hint ? hint : 4 * (len - c.payload > (1u << 20) ? (1u << 20) : len - c.payload) + 65536;
            c.hinted = hint != 0;
        } else {
            c.group = hrc == ZPAQGPU_E_UNSUPPORTED ? -2 : -1;
            c.hint = 0;
        }
        job.cand.push_back(c);
    }
    // The "<N> bytes" comment is untrusted input and a stored block may hold locators of its own: a
    // slot is never larger than what the bytes up to the next candidate can plausibly decode to, and all
    // slots together stay inside a share of the free memory.  An under-estimate only costs the sizing
    // retry of run_decompress (the kernels report the true length).
    {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const u64 budget = std::max<u64>(u64(free_b + ctx->plain.cap) / 2, 1u << 20);
        u64 sum = 0;
        for (size_t i = 0; i < job.cand.size(); ++i) {
            DecCandidate &c = job.cand[i];
            if (c.group < 0) continue;
            const u64 next = i + 1 < job.cand.size() ? job.cand[i + 1].start : len;
            const u64 csz = next > c.payload ? next - c.payload : 0;
            c.hint = std::min<u64>(c.hint, 4096 * csz + (1u << 20));
            sum += c.hint;
        }
        if (sum > budget) {
            const u64 share = std::max<u64>(budget / std::max<size_t>(1, job.cand.size()), 65536);
            for (DecCandidate &c : job.cand)
                if (c.group >= 0) c.hint = std::min<u64>(c.hint, share);
        }
    }
    // 3. decode
    if (!job.cand.empty() && !job.models.empty()) {
        // rejected headers get a dummy group so indices stay aligned; they are never launched
        if ((rc = run_decompress(ctx, job, false))) return rc;
    }
    ctx->stats.h2d_ms = h2d_ms;
    ctx->stats.pack_ms = 0;
    // 4. walk the candidates the way repeated find_block calls would (decompressor.v:219-346)
    u64 pos = 0, total = 0;
    int seg_total = 0, block_index = 0;
    int status = ZPAQGPU_OK;
    size_t rec_at = 0;
    ctx->walk_stopped = false;
    for (size_t i = 0; i < job.cand.size(); ++i) {
        const DecCandidate &c = job.cand[i];
        // a later find_block call restarts its rolling hashes at `pos`, so it only sees locators
        // that lie completely behind the previous block; anything earlier is inside that block
        if (pos > 0 && c.start < pos + 16) continue;
        if (c.group < 0) {
            // find_block returns false here and `for find_block {}` ends (cmd/main.v:349)
            status = c.group == -2 ? ZPAQGPU_E_UNSUPPORTED : ZPAQGPU_OK;
            ctx->walk_stopped = true;
            break;
        }
        while (rec_at < job.recs.size() && job.recs[rec_at].block < u32(i)) ++rec_at;
        size_t r = rec_at;
        for (; r < job.recs.size() && job.recs[r].block == u32(i); ++r) {
            const DecSegRec &rec = job.recs[r];
            DecodedSeg ds;
            ds.seg.block_start = c.start, ds.seg.block_end = c.res.end_pos;
            ds.seg.name_off = rec.name_off, ds.seg.comment_off = rec.comment_off;
            ds.seg.out_off = total, ds.seg.out_len = rec.out_len;
            ds.seg.block_index = block_index, ds.seg.sha1_ok = job.sha_ok[r];
            ds.src = rec.out_off;
            segs.push_back(ds);
            total += rec.out_len;
            ++seg_total;
        }
        if (c.res.status != ZPAQGPU_OK && status == ZPAQGPU_OK) status = c.res.status;
        pos = c.res.end_pos;
        ++block_index;
        if (c.res.status != ZPAQGPU_OK) {
            ctx->walk_stopped = true;
            break;
        }
    }
    (void)seg_total;
    *d_plain = job.d_plain, *status_out = status, *total_out = total;
    return ZPAQGPU_OK;
}

// Plaintext of the listed segments from the device arena to `out`, back to back in list order;
// contiguous runs of the arena are merged into single copies.
int plain_fetch(zpaqgpu_ctx *ctx, const std::vector<DecodedSeg> &list, const u8 *d_plain, uint8_t *out) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[6], st));
    u64 dst = 0;
    for (size_t k = 0; k < list.size();) {
        u64 src = list[k].src, run = list[k].seg.out_len;
        size_t j = k + 1;
        while (j < list.size() && list[j].src == src + run) run += list[j].seg.out_len, ++j;
        if (run) CK(cudaMemcpyAsync(out + dst, d_plain + src, run, cudaMemcpyDeviceToHost, st));
        dst += run;
        k = j;
    }
    CK(cudaEventRecord(ctx->ev[7], st));
    CK(cudaStreamSynchronize(st));
    ctx->stats.d2h_ms = elapsed(ctx->ev[6], ctx->ev[7]);
    return ZPAQGPU_OK;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

const char *zpaqgpu_strerror(int code) {
    switch (code) {
    case ZPAQGPU_OK: return "ok";
    case ZPAQGPU_E_NODEVICE: return "no usable CUDA device (libzpaqgpu has no CPU fallback)";
    case ZPAQGPU_E_CUDA: return "CUDA error";
    case ZPAQGPU_E_NOSPACE: return "output buffer too small";
    case ZPAQGPU_E_ARG: return "bad argument";
    case ZPAQGPU_E_FORMAT: return "malformed archive";
    case ZPAQGPU_E_UNSUPPORTED: return "unsupported model or post-processor";
    case ZPAQGPU_E_STATE: return "call not valid in this state";
    case ZPAQGPU_E_NOMEM: return "out of device memory";
    default: return "unknown error";
    }
}

int zpaqgpu_init(zpaqgpu_ctx **out, int device) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    if (!out) return ZPAQGPU_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return ZPAQGPU_E_NODEVICE;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return ZPAQGPU_E_NODEVICE;
    if (device >= count) return ZPAQGPU_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return ZPAQGPU_E_NODEVICE;
    zpaqgpu_ctx *ctx = new zpaqgpu_ctx();
    ctx->device = device;
    if (const char *v = std::getenv("ZPAQGPU_DECODER"))
        ctx->decoder = std::strcmp(v, "serial") == 0 ? 0 : (std::strcmp(v, "tree") == 0 ? 1 : 2);
    if (const char *v = std::getenv("ZPAQGPU_SPEC_PROBE")) ctx->spec_probe = std::atoi(v) != 0;
    if (const char *v = std::getenv("ZPAQGPU_PULL")) ctx->pull_how = std::atoi(v) & 3;
    if (const char *v = std::getenv("ZPAQGPU_GUESS")) ctx->guess = std::max(0, std::min(4, std::atoi(v)));
    if (const char *v = std::getenv("ZPAQGPU_ENC_FLAGS")) ctx->enc_flags = std::atoi(v) & 3;
    if (const char *v = std::getenv("ZPAQGPU_AHEAD")) ctx->ahead = std::atoi(v) != 0;
    if (const char *v = std::getenv("ZPAQGPU_GENERIC")) ctx->generic_warp = std::strcmp(v, "lane0") != 0;
    if (const char *v = std::getenv("ZPAQGPU_WS_LIMIT_MB")) ctx->ws_limit = u64(std::atoll(v)) << 20;  // profiling: ncu saves
    // and restores every device buffer between replay passes
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) == cudaSuccess;
    ctx->stream = ctx->own_stream;
    for (auto &e : ctx->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming) == cudaSuccess;
    // constant tables: narrowed on the host, uploaded once
    const Tables &T = tables();
    const size_t bytes = 32768 * 2 + 4096 * 2 + 512 + 1024 * 4 + 256 * 4 + 4096 * 2 + 32768 * 2 + 32;
    std::vector<uint8_t> img(bytes);
    int16_t *st = reinterpret_cast<int16_t *>(img.data());
    uint16_t *sq = reinterpret_cast<uint16_t *>(img.data() + 65536);
    uint8_t *nx = img.data() + 65536 + 8192;
    int32_t *dt = reinterpret_cast<int32_t *>(img.data() + 65536 + 8192 + 512);
    int32_t *d2 = dt + 1024;
    for (int i = 0; i < 32768; ++i) st[i] = int16_t(T.stretch[i]);
    for (int i = 0; i < 4096; ++i) sq[i] = uint16_t(T.squash[i]);
    for (int s = 0; s < 256; ++s) nx[s * 2] = T.ns[s * 4], nx[s * 2 + 1] = T.ns[s * 4 + 1];
    std::memcpy(dt, T.dt, sizeof(T.dt));
    std::memcpy(d2, T.dt2k, sizeof(T.dt2k));
    uint16_t *sqp = reinterpret_cast<uint16_t *>(d2 + 256);
    int16_t *stp = reinterpret_cast<int16_t *>(sqp + 4096);
    for (int j = 0; j < 4096; ++j) {  // squash(p) for p = j - 2048: index p+2047 clamped to [0,4093]
        int idx = j - 1;
        idx = idx < 0 ? 0 : (idx > 4093 ? 4093 : idx);
        sqp[j] = uint16_t(T.squash[idx]);
    }
    for (int i = 0; i < 32768; ++i) stp[i] = int16_t(T.stretch[i < 1 ? 1 : i]);
    uint32_t *lk = reinterpret_cast<uint32_t *>(stp + 32768);
    for (int w = 0; w < 8; ++w) lk[w] = 0;
    for (int s = 0; s < 256; ++s)
        if (T.ns[s * 4 + 3] > T.ns[s * 4 + 2]) lk[s >> 5] |= 1u << (s & 31);
    ok = ok && cudaMalloc(&ctx->tables_mem, bytes) == cudaSuccess;
    ok = ok && cudaMemcpy(ctx->tables_mem, img.data(), bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        zpaqgpu_destroy(ctx);
        return ZPAQGPU_E_CUDA;
    }
    u8 *base = static_cast<u8 *>(ctx->tables_mem);
    ctx->tables.stretch = reinterpret_cast<const int16_t *>(base);
    ctx->tables.squash = reinterpret_cast<const u16 *>(base + 65536);
    ctx->tables.nex = base + 65536 + 8192;
    ctx->tables.dt = reinterpret_cast<const i32 *>(base + 65536 + 8192 + 512);
    ctx->tables.dt2k = ctx->tables.dt + 1024;
    ctx->tables.squash_pad = reinterpret_cast<const u16 *>(ctx->tables.dt2k + 256);
    ctx->tables.stretch_pad = reinterpret_cast<const int16_t *>(ctx->tables.squash_pad + 4096);
    ctx->tables.likely = reinterpret_cast<const u32 *>(ctx->tables.stretch_pad + 32768);
    *out = ctx;
    return ZPAQGPU_OK;
    });
}

void zpaqgpu_destroy(zpaqgpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    DevBuf *bufs[] = {&ctx->workspace, &ctx->in, &ctx->arena, &ctx->out, &ctx->desc, &ctx->pay_len,
                      &ctx->digests, &ctx->seg_size, &ctx->out_off, &ctx->modelblob, &ctx->results,
                      &ctx->seg_recs, &ctx->misc, &ctx->heads, &ctx->plain, &ctx->pool,
                      &ctx->jd_in, &ctx->jd_frag, &ctx->jd_tab, &ctx->jd_packed, &ctx->jd_small,
                      &ctx->jd_out2, &ctx->jd_off2};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->tables_mem) cudaFree(ctx->tables_mem);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    delete ctx;
}

const char *zpaqgpu_last_error(const zpaqgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

int zpaqgpu_set_kernel(zpaqgpu_ctx *ctx, int kernel) {
    if (!ctx || kernel < 0 || kernel > 2) return ZPAQGPU_E_ARG;
    ctx->kernel_pref = kernel;
    return ZPAQGPU_OK;
}
int zpaqgpu_set_table_mode(zpaqgpu_ctx *ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return ZPAQGPU_E_ARG;
    ctx->table_mode = mode;
    return ZPAQGPU_OK;
}
int zpaqgpu_set_workspace_limit(zpaqgpu_ctx *ctx, uint64_t bytes) {
    if (!ctx) return ZPAQGPU_E_ARG;
    ctx->ws_limit = bytes;
    return ZPAQGPU_OK;
}
int zpaqgpu_set_stream(zpaqgpu_ctx *ctx, void *cuda_stream) {
    if (!ctx) return ZPAQGPU_E_ARG;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return ZPAQGPU_OK;
}

int zpaqgpu_level_header(int level, uint8_t *out, int cap) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    const std::vector<uint8_t> h = level_header(level);
    if (int(h.size()) > cap || !out) return ZPAQGPU_E_NOSPACE;
    std::memcpy(out, h.data(), h.size());
    return int(h.size());
    });
}

int zpaqgpu_tables(int32_t *squash4096, int32_t *stretch32768, uint8_t *state1024) {
    const Tables &T = tables();
    if (squash4096) std::memcpy(squash4096, T.squash, sizeof(T.squash));
    if (stretch32768) std::memcpy(stretch32768, T.stretch, sizeof(T.stretch));
    if (state1024) std::memcpy(state1024, T.ns, sizeof(T.ns));
    return ZPAQGPU_OK;
}

int zpaqgpu_describe_model(const uint8_t *header, int header_len, zpaqgpu_model_info *out) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    if (!out) return ZPAQGPU_E_ARG;
    Model m;
    const int rc = model_from_level_layout(header, header_len, m);
    if (rc) return rc;
    out->n = m.n, out->cend = m.cend, out->hbegin = m.hbegin, out->hend = m.hend;
    out->hsize = (m.cend + 1) + (m.hend - m.hbegin + 1);
    out->is_chain = m.is_chain, out->n_isse = m.n_isse, out->has_mix2 = m.has_mix2;
    out->ctx_mode = m.ctx_mode, out->n_hash = m.n_hash;
    out->workspace_bytes = m.ws_bytes;
    out->hash_table_bytes = 0;
    for (const CompDesc &c : m.comps)
        if (c.type == C_ICM || c.type == C_ISSE) out->hash_table_bytes += c.ht_len;
    return ZPAQGPU_OK;
    });
}

int zpaqgpu_last_stats(const zpaqgpu_ctx *ctx, zpaqgpu_stats *out) {
    if (!ctx || !out) return ZPAQGPU_E_ARG;
    *out = ctx->stats;
    return ZPAQGPU_OK;
}

// ---- batch compression ----
int zpaqgpu_compress_blocks_header(zpaqgpu_ctx *ctx, const uint8_t *header, int header_len, const uint8_t *in,
                                   const uint64_t *in_off, int n_blocks, const char *const *names,
                                   const char *const *comments, uint8_t *out, uint64_t out_cap,
                                   uint64_t *out_off, uint64_t *out_need) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx) return ZPAQGPU_E_ARG;
    Model m;
    const int rc = model_from_level_layout(header, header_len, m);
    if (rc) {
        ctx->err = m.error;
        return rc;
    }
    return compress_host(ctx, m, in, in_off, n_blocks, names, comments, out, out_cap, out_off, out_need);
    });
}

int zpaqgpu_compress_blocks(zpaqgpu_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off, int n_blocks,
                            const char *const *names, const char *const *comments, uint8_t *out,
                            uint64_t out_cap, uint64_t *out_off, uint64_t *out_need) {
    return zg::guarded<int>(ctx, [&]() -> int {
    const std::vector<uint8_t> h = level_header(level);
    return zpaqgpu_compress_blocks_header(ctx, h.data(), int(h.size()), in, in_off, n_blocks, names, comments,
                                          out, out_cap, out_off, out_need);
    });
}

int zpaqgpu_compress_blocks_dev(zpaqgpu_ctx *ctx, int level, const void *d_in, const void *d_in_off,
                                const uint64_t *h_in_off, int n_blocks, void *d_out, uint64_t out_cap,
                                void *d_out_off, uint64_t *out_total) {
    return zg::guarded<int>(ctx, [&]() -> int {
    (void)d_in_off;
    if (!ctx || n_blocks <= 0 || !h_in_off || !d_out || !d_out_off) return ZPAQGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const std::vector<uint8_t> h = level_header(level);
    Model m;
    int rc = model_from_level_layout(h.data(), int(h.size()), m);
    if (rc) return ctx->err = m.error, rc;
    CompressJob job;
    job.model = &m;
    job.blocks.resize(size_t(n_blocks));
    job.segs.resize(size_t(n_blocks));
    std::vector<std::string> comments(static_cast<size_t>(n_blocks));
    for (int b = 0; b < n_blocks; ++b) {
        job.blocks[size_t(b)] = EncBlock{u32(b), 1};
        SegSpec &sp = job.segs[size_t(b)];
        sp.in_off = h_in_off[b], sp.in_len = h_in_off[b + 1] - h_in_off[b];
        comments[size_t(b)] = size_comment(sp.in_len);  // cmd/main.v:303
        sp.name = "", sp.comment = comments[size_t(b)].c_str();
        sp.called = true;
    }
    job.d_in = static_cast<const u8 *>(d_in);
    job.d_out = static_cast<u8 *>(d_out), job.out_cap = out_cap;
    job.d_out_off = static_cast<u64 *>(d_out_off);
    if ((rc = run_compress(ctx, job))) return rc;
    if (out_total) *out_total = job.total;
    return job.fits ? ZPAQGPU_OK : ZPAQGPU_E_NOSPACE;
    });
}

// ---- locator scan ----
int zpaqgpu_find_blocks(zpaqgpu_ctx *ctx, const uint8_t *arc, uint64_t len, uint64_t *starts, int cap,
                        int *n_found) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || (len && !arc) || cap < 0 || !n_found) return ZPAQGPU_E_ARG;
    *n_found = 0;
    if (len == 0) return ZPAQGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ensure(ctx, ctx->in, len))) return rc;
    const u32 dcap = u32(std::max(cap, 1024));
    if ((rc = ensure(ctx, ctx->results, 8 * size_t(dcap)))) return rc;
    if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
    CK(cudaMemcpyAsync(ctx->in.p, arc, len, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
    launch_find_blocks(static_cast<const u8 *>(ctx->in.p), len, static_cast<u64 *>(ctx->results.p), dcap,
                       static_cast<u32 *>(ctx->misc.p), st);
    CK(cudaGetLastError());
    u32 count = 0;
    CK(cudaMemcpyAsync(&count, ctx->misc.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_found = int(count);
    if (count > u32(cap)) return ZPAQGPU_E_NOSPACE;
    std::vector<u64> tmp(count);
    if (count) CK(cudaMemcpy(tmp.data(), ctx->results.p, 8 * size_t(count), cudaMemcpyDeviceToHost));
    std::sort(tmp.begin(), tmp.end());
    if (count) std::memcpy(starts, tmp.data(), 8 * size_t(count));
    return ZPAQGPU_OK;
    });
}

// ---- whole-archive decompression ----
int zpaqgpu_decompress_archive(zpaqgpu_ctx *ctx, const uint8_t *arc, uint64_t len, uint8_t *out, uint64_t out_cap,
                               uint64_t *out_need, zpaqgpu_segment *segs, int segs_cap, int *n_segs) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || (len && !arc)) return ZPAQGPU_E_ARG;
    if (out_need) *out_need = 0;
    if (n_segs) *n_segs = 0;
    if (len == 0) return ZPAQGPU_OK;
    std::vector<DecodedSeg> list;
    const u8 *d_plain = nullptr;
    int status = ZPAQGPU_OK;
    u64 total = 0;
    int rc = decode_archive_dev(ctx, arc, len, list, &d_plain, &status, &total);
    if (rc) return rc;
    const int seg_total = int(list.size());
    if (segs)
        for (int k = 0; k < seg_total && k < segs_cap; ++k) segs[k] = list[size_t(k)].seg;
    if (out_need) *out_need = total;
    if (n_segs) *n_segs = seg_total;
    if (total > out_cap || (segs && seg_total > segs_cap)) return ZPAQGPU_E_NOSPACE;
    if (total && !out) return ZPAQGPU_E_ARG;
    if ((rc = plain_fetch(ctx, list, d_plain, out))) return rc;
    return status;
    });
}

int zpaqgpu_decompress_blocks_dev(zpaqgpu_ctx *ctx, const void *d_arc, const uint64_t *h_arc_off, int n_blocks,
                                  void *d_out, const uint64_t *h_out_off, void *d_out_len, int *n_bad) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || n_blocks <= 0 || !d_arc || !h_arc_off || !d_out || !h_out_off) return ZPAQGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    // block heads to the host for header parsing
    const u32 head = 512;
    if ((rc = ensure(ctx, ctx->heads, size_t(head) * size_t(n_blocks) + 8 * size_t(n_blocks + 1)))) return rc;
    u64 *d_off = reinterpret_cast<u64 *>(static_cast<u8 *>(ctx->heads.p) + size_t(head) * size_t(n_blocks));
    d_off = reinterpret_cast<u64 *>(align_up(reinterpret_cast<uintptr_t>(d_off), 8));
    if ((rc = ensure(ctx, ctx->out_off, 8 * size_t(n_blocks + 1)))) return rc;
    CK(cudaMemcpyAsync(ctx->out_off.p, h_arc_off, 8 * size_t(n_blocks + 1), cudaMemcpyHostToDevice, st));
    launch_gather_heads(static_cast<const u8 *>(d_arc), static_cast<const u64 *>(ctx->out_off.p), n_blocks, head,
                        static_cast<u8 *>(ctx->heads.p), st);
    CK(cudaGetLastError());
    std::vector<uint8_t> heads(size_t(head) * size_t(n_blocks));
    CK(cudaMemcpyAsync(heads.data(), ctx->heads.p, heads.size(), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    DecompressJob job;
    job.d_arc = static_cast<const u8 *>(d_arc), job.arc_len = h_arc_off[n_blocks];
    job.d_plain = static_cast<u8 *>(d_out);
    std::map<std::vector<uint8_t>, int> group_of;
    for (int b = 0; b < n_blocks; ++b) {
        const uint8_t *hp = heads.data() + size_t(head) * size_t(b);
        const u64 avail = std::min<u64>(head, h_arc_off[b + 1] - h_arc_off[b]);
        DecCandidate c;
        if (avail < 16) return ctx->err = "block shorter than its locator", ZPAQGPU_E_FORMAT;
        c.start = h_arc_off[b] + 16;
        Model m;
        u64 used = 0;
        const int hrc = model_from_archive(hp + 16, avail - 16, m, &used);
        if (hrc != ZPAQGPU_OK) return ctx->err = "block " + std::to_string(b) + ": " + m.error, hrc;
        c.payload = c.start + used;
        auto it = group_of.find(m.header);
        if (it == group_of.end()) {
            it = group_of.emplace(m.header, int(job.models.size())).first;
            job.models.push_back(m);
        }
        c.group = it->second;
        c.out_off = h_out_off[b];
        c.hint = h_out_off[b + 1] - h_out_off[b];
        job.cand.push_back(c);
    }
    if ((rc = run_decompress(ctx, job, true))) return rc;
    int bad = 0;
    std::vector<u64> lens(static_cast<size_t>(n_blocks));
    for (int b = 0; b < n_blocks; ++b) {
        lens[size_t(b)] = job.cand[size_t(b)].res.out_len;
        if (job.cand[size_t(b)].res.status != ZPAQGPU_OK) ++bad;
    }
    for (size_t r = 0; r < job.recs.size(); ++r)
        if (job.sha_ok[r] == 0) ++bad;
    if (d_out_len) CK(cudaMemcpy(d_out_len, lens.data(), 8 * size_t(n_blocks), cudaMemcpyHostToDevice));
    if (n_bad) *n_bad = bad;
    return ZPAQGPU_OK;
    });
}

}  // extern "C"

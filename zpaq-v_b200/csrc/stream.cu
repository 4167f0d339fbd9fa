// stream.cu -- the streaming-shaped calls of the C ABI: what the V/C++/Python mirrors of
// zpaq.Compressor call method by method (compressor.v:79-413).  Bytes are gathered per segment on the
// host; a block is coded on the device either at once (zpaqgpu_block_end) or, for callers that write
// many blocks one after the other like cmd/main.v:288-317 (one block per file), queued at
// zpaqgpu_block_end_queue and coded together at zpaqgpu_flush -- one launch with a chain per block
// instead of one launch per block with a single chain on 148 SMs.  Output order is queue order.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.h"

using namespace zg;

namespace {

// Codes the blocks `q` (queue order) and delivers their bytes back to back: into `out` when they fit
// `cap`, else into `spill` (ZPAQGPU_E_NOSPACE, *need = size).  Blocks of one model form one job.
int code_queue(zpaqgpu_ctx *ctx, const std::vector<QueuedBlock> &q, uint8_t *out, u64 cap, u64 *need,
               std::vector<uint8_t> &spill) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nq = q.size();
    std::vector<int> groups;  // model indices in order of first appearance
    for (const QueuedBlock &b : q)
        if (std::find(groups.begin(), groups.end(), b.model) == groups.end()) groups.push_back(b.model);
    bool any_empty = false;
    for (const QueuedBlock &b : q) any_empty = any_empty || b.segs.empty();
    const bool direct = groups.size() == 1 && !any_empty;  // the device archive IS the result
    std::vector<std::vector<uint8_t>> parts(direct ? 0 : nq);  // else: every block's bytes on the host
    zpaqgpu_stats sum{};
    for (int g : groups) {
        const Model &m = ctx->st_models[size_t(g)];
        CompressJob job;
        job.model = &m;
        std::vector<size_t> members;
        u64 total_in = 0, worst = 64;
        for (size_t k = 0; k < nq; ++k) {
            if (q[k].model != g) continue;
            if (q[k].segs.empty()) {
                // a block without segments is just its header and 0xFF (compressor.v:150-181, :409)
                parts[k] = m.block_prefix;
                parts[k].push_back(0xFF);
                continue;
            }
            members.push_back(k);
            job.blocks.push_back(EncBlock{u32(job.segs.size()), u32(q[k].segs.size())});
            worst += m.block_prefix.size() + 8;
            for (const PendingSeg &s : q[k].segs) {
                SegSpec sp;
                sp.name = s.name.c_str(), sp.comment = s.comment.c_str();
                sp.in_off = total_in, sp.in_len = s.data.size(), sp.called = s.called;
                job.segs.push_back(sp);
                total_in += sp.in_len;
                worst += sp.in_len + sp.in_len / 4 + 2048 + s.name.size() + s.comment.size();
            }
        }
        if (members.empty()) continue;
        const int nb = int(members.size());
        int rc;
        if ((rc = ensure(ctx, ctx->in, std::max<u64>(total_in, 16)))) return rc;
        CK(cudaEventRecord(ctx->ev[6], st));
        {   // every segment straight from its host vector to its place on the device
            size_t si = 0;
            for (size_t k : members)
                for (const PendingSeg &s : q[k].segs) {
                    if (!s.data.empty())
                        CK(cudaMemcpyAsync(static_cast<u8 *>(ctx->in.p) + job.segs[si].in_off, s.data.data(), s.data.size(),
                                           cudaMemcpyHostToDevice, st));
                    ++si;
                }
        }
        CK(cudaEventRecord(ctx->ev[7], st));
        CK(cudaStreamSynchronize(st));
        const float h2d_ms = elapsed(ctx->ev[6], ctx->ev[7]);
        for (int attempt = 0;; ++attempt) {
            if ((rc = ensure(ctx, ctx->out, worst))) return rc;
            if ((rc = ensure(ctx, ctx->out_off, 8 * size_t(nb + 1)))) return rc;
            job.d_in = static_cast<const u8 *>(ctx->in.p);
            job.d_out = static_cast<u8 *>(ctx->out.p), job.out_cap = ctx->out.cap;
            job.d_out_off = static_cast<u64 *>(ctx->out_off.p);
            if ((rc = run_compress(ctx, job))) return rc;
            if (job.fits) break;
            if (attempt == 1) return ZPAQGPU_E_NOSPACE;
            worst = job.total;
        }
        ctx->stats.h2d_ms = h2d_ms;
        sum.init_ms += ctx->stats.init_ms, sum.codec_ms += ctx->stats.codec_ms, sum.sha1_ms += ctx->stats.sha1_ms;
        sum.pack_ms += ctx->stats.pack_ms, sum.h2d_ms += h2d_ms;
        sum.launches += ctx->stats.launches, sum.codec_launches += ctx->stats.codec_launches;
        sum.waves += ctx->stats.waves, sum.retries += ctx->stats.retries;
        sum.kernel = ctx->stats.kernel, sum.warps_per_cta = ctx->stats.warps_per_cta;
        sum.workspace_bytes_per_block = ctx->stats.workspace_bytes_per_block;
        sum.pool_bytes_used = std::max(sum.pool_bytes_used, ctx->stats.pool_bytes_used), sum.paged = ctx->stats.paged;
        if (direct) {
            if (need) *need = job.total;
            uint8_t *dst = out;
            if (job.total > cap || (!out && job.total)) {
                spill.resize(size_t(job.total));
                dst = spill.data();
            }
            CK(cudaEventRecord(ctx->ev[6], st));
            if (job.total) CK(cudaMemcpyAsync(dst, ctx->out.p, job.total, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ctx->ev[7], st));
            CK(cudaStreamSynchronize(st));
            sum.d2h_ms = elapsed(ctx->ev[6], ctx->ev[7]);
            ctx->stats = sum;
            return dst == out ? ZPAQGPU_OK : ZPAQGPU_E_NOSPACE;
        }
        std::vector<u64> off(size_t(nb) + 1);
        std::vector<uint8_t> arc(static_cast<size_t>(job.total));
        CK(cudaMemcpyAsync(off.data(), ctx->out_off.p, 8 * off.size(), cudaMemcpyDeviceToHost, st));
        if (job.total) CK(cudaMemcpyAsync(arc.data(), ctx->out.p, job.total, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int b = 0; b < nb; ++b) parts[members[size_t(b)]].assign(arc.begin() + off[size_t(b)], arc.begin() + off[size_t(b) + 1]);
    }
    ctx->stats = sum;
    u64 total = 0;
    for (const auto &p : parts) total += p.size();
    if (need) *need = total;
    uint8_t *dst = out;
    if (total > cap || (!out && total)) {
        spill.resize(size_t(total));
        dst = spill.data();
    }
    u64 at = 0;
    for (const auto &p : parts) {
        if (!p.empty()) std::memcpy(dst + at, p.data(), p.size());
        at += p.size();
    }
    return dst == out ? ZPAQGPU_OK : ZPAQGPU_E_NOSPACE;
}

// Finished bytes kept from a call whose buffer was too small.
int64_t serve_done(zpaqgpu_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *need) {
    if (need) *need = ctx->st_done.size();
    if (ctx->st_done.size() > cap || (!out && !ctx->st_done.empty())) return ZPAQGPU_E_NOSPACE;
    if (!ctx->st_done.empty()) std::memcpy(out, ctx->st_done.data(), ctx->st_done.size());
    const int64_t n = int64_t(ctx->st_done.size());
    ctx->st_has_done = false;
    std::vector<uint8_t>().swap(ctx->st_done);
    return n;
}

void queue_clear(zpaqgpu_ctx *ctx) {
    ctx->st_queue.clear();
    ctx->st_queue_bytes = 0;
    if (ctx->st_state == 2) ctx->st_models.clear(), ctx->st_model = -1;
}

}  // namespace

extern "C" {

int zpaqgpu_block_begin_header(zpaqgpu_ctx *ctx, const uint8_t *header, int header_len) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (ctx->st_state != 2 || ctx->st_has_done) return ZPAQGPU_E_STATE;  // compressor.v:80-82
    Model m;
    const int rc = model_from_level_layout(header, header_len, m);
    if (rc) return ctx->err = m.error, rc;
    int idx = -1;
    for (size_t k = 0; k < ctx->st_models.size(); ++k)
        if (ctx->st_models[k].header == m.header) idx = int(k);
    if (idx < 0) {
        ctx->st_models.push_back(std::move(m));
        idx = int(ctx->st_models.size()) - 1;
    }
    ctx->st_model = idx;
    ctx->st_segs.clear();
    ctx->st_state = 0;
    return ZPAQGPU_OK;
    });
}
int zpaqgpu_block_begin(zpaqgpu_ctx *ctx, int level) {
    return zg::guarded<int>(ctx, [&]() -> int {
    const std::vector<uint8_t> h = level_header(level);
    return zpaqgpu_block_begin_header(ctx, h.data(), int(h.size()));
    });
}
int zpaqgpu_segment_begin(zpaqgpu_ctx *ctx, const char *filename, const char *comment) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (ctx->st_state != 0) return ZPAQGPU_E_STATE;  // compressor.v:213-215
    PendingSeg s;
    s.name = filename ? filename : "", s.comment = comment ? comment : "";
    ctx->st_segs.push_back(std::move(s));
    ctx->st_state = 1;
    return ZPAQGPU_OK;
    });
}
int zpaqgpu_segment_write(zpaqgpu_ctx *ctx, const uint8_t *data, uint64_t len) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx || (len && !data)) return ZPAQGPU_E_ARG;
    if (ctx->st_state != 1) return ZPAQGPU_E_STATE;  // compressor.v:260-262
    PendingSeg &s = ctx->st_segs.back();
    s.called = true;
    s.data.insert(s.data.end(), data, data + len);
    return ZPAQGPU_OK;
    });
}
int zpaqgpu_segment_end(zpaqgpu_ctx *ctx) {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (ctx->st_state != 1) return ZPAQGPU_E_STATE;  // compressor.v:358-360
    ctx->st_state = 0;
    return ZPAQGPU_OK;
}

int64_t zpaqgpu_block_end(zpaqgpu_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *need) {
    return zg::guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (ctx->st_has_done) return serve_done(ctx, out, cap, need);
    if (ctx->st_state != 0) return ZPAQGPU_E_STATE;  // compressor.v:403-405
    if (!ctx->st_queue.empty()) {
        ctx->err = "blocks are queued: zpaqgpu_flush must deliver them before a block is coded on its own";
        return ZPAQGPU_E_STATE;
    }
    std::vector<QueuedBlock> one(1);
    one[0].model = ctx->st_model;
    one[0].segs = std::move(ctx->st_segs);
    ctx->st_segs.clear();
    u64 n = 0;
    const int rc = code_queue(ctx, one, out, cap, &n, ctx->st_done);
    if (rc != ZPAQGPU_OK && rc != ZPAQGPU_E_NOSPACE) {
        ctx->st_segs = std::move(one[0].segs);  // nothing was delivered: the block can be ended again
        return rc;
    }
    ctx->st_state = 2;
    queue_clear(ctx);
    if (need) *need = n;
    if (rc == ZPAQGPU_E_NOSPACE) {
        ctx->st_has_done = true;  // kept: call again with a larger buffer
        return ZPAQGPU_E_NOSPACE;
    }
    return int64_t(n);
    });
}

int zpaqgpu_stream_batch(zpaqgpu_ctx *ctx, int max_blocks, uint64_t max_bytes) {
    if (!ctx || max_blocks < 0) return ZPAQGPU_E_ARG;
    ctx->st_batch_blocks = max_blocks ? max_blocks : 1024;
    ctx->st_batch_bytes = max_bytes ? max_bytes : (1ull << 30);
    return ZPAQGPU_OK;
}

int zpaqgpu_block_end_queue(zpaqgpu_ctx *ctx) {
    return zg::guarded<int>(ctx, [&]() -> int {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (ctx->st_state != 0 || ctx->st_has_done) return ZPAQGPU_E_STATE;  // compressor.v:403-405
    QueuedBlock b;
    b.model = ctx->st_model;
    b.segs = std::move(ctx->st_segs);
    ctx->st_segs.clear();
    for (const PendingSeg &s : b.segs) ctx->st_queue_bytes += s.data.size();
    ctx->st_queue.push_back(std::move(b));
    ctx->st_state = 2;
    return (int(ctx->st_queue.size()) >= ctx->st_batch_blocks || ctx->st_queue_bytes >= ctx->st_batch_bytes) ? 1 : 0;
    });
}

int zpaqgpu_queued(const zpaqgpu_ctx *ctx, int *n_blocks, uint64_t *in_bytes) {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (n_blocks) *n_blocks = int(ctx->st_queue.size());
    if (in_bytes) *in_bytes = ctx->st_queue_bytes;
    return ZPAQGPU_OK;
}

int64_t zpaqgpu_flush(zpaqgpu_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *need) {
    return zg::guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return ZPAQGPU_E_ARG;
    if (ctx->st_has_done) return serve_done(ctx, out, cap, need);
    if (ctx->st_state != 2) return ZPAQGPU_E_STATE;  // a block is still open
    if (need) *need = 0;
    if (ctx->st_queue.empty()) return 0;
    u64 n = 0;
    const int rc = code_queue(ctx, ctx->st_queue, out, cap, &n, ctx->st_done);
    if (rc != ZPAQGPU_OK && rc != ZPAQGPU_E_NOSPACE) return rc;  // the queue is kept
    queue_clear(ctx);
    if (need) *need = n;
    if (rc == ZPAQGPU_E_NOSPACE) {
        ctx->st_has_done = true;
        return ZPAQGPU_E_NOSPACE;
    }
    return int64_t(n);
    });
}

}  // extern "C"

// ctx.h -- internals of libzpaqgpu shared by api.cu and jidac.cu: the context, grow-only device
// buffers and the compression job that both the block API and the jidac front end submit.
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <exception>
#include <new>
#include <string>
#include <vector>

#include "../../include/zpaqgpu.h"
#include "common.cuh"
#include "kernels.h"
#include "model.h"

namespace zg {
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct PendingSeg {
    std::string name, comment;
    std::vector<uint8_t> data;
    bool called = false;  // compress() was called at least once
};

struct QueuedBlock {  // a finished block of the streaming-shaped calls
    int model;            // index into zpaqgpu_ctx::st_models
    std::vector<PendingSeg> segs;
};

struct SegSpec {  // one segment of a compression job
    const char *name, *comment;
    u64 in_off, in_len;
    bool called;
};

}  // namespace zg

struct zpaqgpu_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_side = nullptr, ev_main = nullptr;
    zg::DevTables tables{};
    void *tables_mem = nullptr;
    int kernel_pref = ZPAQGPU_KERNEL_AUTO;
    int table_mode = ZPAQGPU_TABLES_AUTO;
    int enc_flags = 1;         // ZPAQGPU_ENC_FLAGS: 0 no L1 pull in the encoder's history warp, 1 pull the next nibble's slot
                               // line (default), 3 pull the line of the nibble after next (two nibbles of lead)
    int pull_how = 0;          // ZPAQGPU_PULL: how (experiments)
    int guess = 1;             // ZPAQGPU_GUESS=n: the two-warp decoder pulls the n (0, 1, 2, 4) likeliest next slot lines a nibble early
    bool generic_warp = true;  // ZPAQGPU_GENERIC=lane0: the one-lane generic kernels also for headers the warp kernel takes (A/B)
    bool ahead = true;         // ZPAQGPU_AHEAD=0: the encoder on paged tables reads page-table entries when it probes (A/B against the PAGED variant)
    bool spec_probe = true;    // ZPAQGPU_SPEC_PROBE=0: the chain decoders probe only once a nibble is complete
    int decoder = 1;           // ZPAQGPU_DECODER in the environment: serial (0, one bit at a time), tree (1, one warp per
                               // block; the default), tree2 (2, two warps per block where the model allows it; opt-in)
    zg::u64 ws_limit = 0;
    int sm_count = 148;
    std::string err;
    bool walk_stopped = false;  // the last archive walk ended before its last candidate (find_block false / bad block)
    zpaqgpu_stats stats{};
    // grow-only device buffers
    zg::DevBuf workspace, in, arena, out, desc, pay_len, digests, seg_size, out_off, modelblob, results,
        seg_recs, misc, heads, plain, pool;
    // jidac front end (jidac.cu): input, fragment tables, dedup table, packed stored fragments,
    // index-block plaintext and a second archive buffer for the c/h/i blocks
    zg::DevBuf jd_in, jd_frag, jd_tab, jd_packed, jd_small, jd_out2, jd_off2;
    zpaqgpu_jidac_stats jd_stats{};
    // pinned host staging for small read-backs
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    // streaming-shaped state (compressor.v:6-8 state machine), stream.cu
    int st_state = 2;  // 0 block, 1 segment, 2 start
    int st_model = -1; // index into st_models of the block being filled
    std::vector<zg::Model> st_models;        // distinct model headers of the queued blocks
    std::vector<zg::PendingSeg> st_segs;     // segments of the block being filled
    std::vector<zg::QueuedBlock> st_queue;   // finished blocks waiting for zpaqgpu_flush
    zg::u64 st_queue_bytes = 0;
    int st_batch_blocks = 1024;              // block_end_queue reports "full" at these limits
    zg::u64 st_batch_bytes = 1ull << 30;
    std::vector<uint8_t> st_done;  // coded bytes kept until the caller's buffer is large enough
    bool st_has_done = false;
};


#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return ZPAQGPU_E_CUDA;                                                            \
        }                                                                                     \
    } while (0)

namespace zg {

// NVTX range around a stage of a call (tracing): visible in nsys / ncu --nvtx, free otherwise.
struct Range {
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range &) = delete;
    Range &operator=(const Range &) = delete;
};

// Nothing may leave the C ABI as a C++ exception: host allocation failures and the like become status codes.
template <class R, class F>
R guarded(zpaqgpu_ctx *ctx, F &&f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        if (ctx) ctx->err = "host allocation failed";
        return R(ZPAQGPU_E_NOMEM);
    } catch (const std::exception &e) {
        if (ctx) ctx->err = e.what();
        return R(ZPAQGPU_E_CUDA);
    } catch (...) {
        if (ctx) ctx->err = "unknown exception";
        return R(ZPAQGPU_E_CUDA);
    }
}

struct CompressJob {
    const Model *model;
    std::vector<EncBlock> blocks;
    std::vector<SegSpec> segs;
    const u8 *d_in;       // plaintext on the device
    u8 *d_out;            // where the archive bytes go (device)
    u64 out_cap;
    u64 *d_out_off;       // device, n_blocks+1
    // results
    u64 total = 0;
    bool fits = true;
};

int ensure(zpaqgpu_ctx *ctx, DevBuf &b, size_t bytes);
int ensure_pinned(zpaqgpu_ctx *ctx, size_t bytes);
u64 align_up(u64 v, u64 a);
float elapsed(cudaEvent_t a, cudaEvent_t b);
// blocks of segments, plaintext already on the device -> archive bytes on the device
int run_compress(zpaqgpu_ctx *ctx, CompressJob &job);

// Host buffers -> archive in ctx->out / ctx->out_off on the device (compress_stage), and the copy back
// (compress_fetch; `shift` is added to every offset).  zpaqgpu_compress_blocks is one after the other; the
// multi-device calls (multi.cu) run all stages first and fetch once every device's size is known.
int compress_stage(zpaqgpu_ctx *ctx, const Model &m, const uint8_t *in, const uint64_t *in_off, int n_blocks,
                   const char *const *names, const char *const *comments, u64 *total);
int compress_fetch(zpaqgpu_ctx *ctx, int n_blocks, u64 total, uint8_t *out, uint64_t *out_off, u64 shift);

// A decoded segment: the public record plus where its plaintext sits in the device arena.
struct DecodedSeg {
    zpaqgpu_segment seg;
    u64 src;
};
// Whole archive (host bytes) -> plaintext of every segment in the device arena, segments in the order
// find_block / find_filename meet them (decompressor.v:219-635).
int decode_archive_dev(zpaqgpu_ctx *ctx, const uint8_t *arc, u64 len, std::vector<DecodedSeg> &segs,
                       const u8 **d_plain, int *status, u64 *total);

// plaintext of the listed segments, device arena -> out, back to back in list order
int plain_fetch(zpaqgpu_ctx *ctx, const std::vector<DecodedSeg> &list, const u8 *d_plain, uint8_t *out);

}  // namespace zg

// ctx.h -- internals of libzpaqgpu shared by api.cu and jidac.cu: the context, grow-only device
// buffers and the compression job that both the block API and the jidac front end submit.
#pragma once
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
    bool enc_l1_pull = true;   // ZPAQGPU_ENC_FLAGS=0: the encoder's history warp does not pull the next slot line into L1
    int pull_how = 0;          // ZPAQGPU_PULL: how (experiments)
    int guess = 1;             // ZPAQGPU_GUESS=n: the two-warp decoder pulls the n (0, 1, 2, 4) likeliest next slot lines a nibble early
    bool spec_probe = true;    // ZPAQGPU_SPEC_PROBE=0: the chain decoders probe only once a nibble is complete
    int decoder = 1;           // ZPAQGPU_DECODER in the environment: serial (0, one bit at a time), tree (1, one warp per
                               // block), tree2 (2, two warps per block; the default where the model allows it)
    zg::u64 ws_limit = 0;
    int sm_count = 148;
    std::string err;
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
    // streaming-shaped state (compressor.v:6-8 state machine)
    int st_state = 2;  // 0 block, 1 segment, 2 start
    zg::Model st_model;
    std::vector<zg::PendingSeg> st_segs;
    std::vector<uint8_t> st_done;  // finished block kept until the caller's buffer is large enough
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

// A decoded segment: the public record plus where its plaintext sits in the device arena.
struct DecodedSeg {
    zpaqgpu_segment seg;
    u64 src;
};
// Whole archive (host bytes) -> plaintext of every segment in the device arena, segments in the order
// find_block / find_filename meet them (decompressor.v:219-635).
int decode_archive_dev(zpaqgpu_ctx *ctx, const uint8_t *arc, u64 len, std::vector<DecodedSeg> &segs,
                       const u8 **d_plain, int *status, u64 *total);

}  // namespace zg

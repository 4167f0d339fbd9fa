// kernels.h -- kernel argument blocks and host-callable launchers of libzpaqgpu.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace zg {

struct EncodeArgs {
    ModelDev model;
    DevTables tables;
    u8 *workspace;           // n_blocks slots of model.ws_bytes
    const u8 *in;            // plaintext
    u8 *arena;               // coded payload slots
    const EncBlock *blocks;  // all blocks of the call
    const EncSeg *segs;
    u64 *pay_len;            // per segment: coded payload bytes (counted even past pay_cap)
    const u32 *order;        // dispatch order: slot k of the wave codes block order[first_block + k]
                             // (largest blocks first, so that CTAs hold blocks of similar length)
    int first_block;         // wave offset into order[]
    int n_blocks;            // blocks in this wave
    int flags;               // 1: the history warp pulls the next nibble's slot line into L1 one step early
};

struct DecodeArgs {
    ModelDev model;
    DevTables tables;
    u8 *workspace;
    const u8 *arc;
    u64 arc_len;
    u8 *out;
    const DecBlock *blocks;
    DecBlockOut *results;
    DecSegRec *seg_recs;
    u32 *seg_count;
    u32 seg_cap;
    const u32 *order;        // dispatch order, as in EncodeArgs
    int first_block;
    int n_blocks;
    int flags;               // tree decoder: 1 = request the next nibble's candidate slots two bits early
};

// generic (all nine component types, ZPAQL interpreter)
__global__ void k_encode_generic(EncodeArgs A);
__global__ void k_decode_generic(DecodeArgs A);

// warp program for any header with up to 32 components (kernels_genwarp.cu): component i on lane i,
// warp-uniform ZPAQL interpreter.  The one-lane kernels above remain for headers with more components.
bool genwarp_supports(const Model &m);
bool launch_encode_genwarp(const EncodeArgs &A, cudaStream_t s);
bool launch_decode_genwarp(const DecodeArgs &A, cudaStream_t s);

// specialised ICM + ISSE chain (+ MIX2): return false when no instantiation fits the model.
// Encoder (kernels_encpipe.cu): three warps per block; decoder (kernels_chain.cu): one warp per block.
bool launch_encode_pipe3(const Model &m, const EncodeArgs &A, int blocks_per_cta, cudaStream_t s);
size_t encpipe_smem_bytes(const Model &m, int blocks_per_cta);
int encpipe_max_blocks_per_cta(const Model &m);
// tree: speculative whole-nibble decoder (one lane per tree node and outcome) instead of the serial one
bool launch_decode_chain(const Model &m, const DecodeArgs &A, int warps_per_cta, bool tree, cudaStream_t s);
size_t chain_smem_bytes(const Model &m, int warps_per_cta);
int chain_max_warps_per_cta(const Model &m);
// Two-warp tree decoder (kernels_dectree.cu): ICM + up to four ISSEs, no MIX2.
bool tree2_supports(const Model &m);
bool launch_decode_tree2(const Model &m, const DecodeArgs &A, int pairs_per_cta, cudaStream_t s);
size_t tree2_smem_bytes(const Model &m, int pairs_per_cta);
int tree2_max_pairs_per_cta(const Model &m);

// ---- auxiliary kernels (kernels_aux.cu) ----
struct ShaJob {
    u64 off, len;  // byte range inside `base`
};
// k_sha1_segments: one byte range per lane, staged through shared memory; digests[j*20..] (sha1.v:42-146)
void launch_sha1(const u8 *base, const ShaJob *jobs, int n_jobs, u8 *digests, cudaStream_t s);

// k_fill_workspace: writes the non-zero initial table images into every workspace slot
struct FillArgs {
    u8 *workspace;
    u64 ws_bytes;
    int n_slots;
    const FillRegion *regions;
    int n_regions;
    const u32 *image;
};
void launch_fill(const FillArgs &A, cudaStream_t s);

// Block assembly for compression (compressor.v:150-181, :217-235, :380-395, :409).
struct PackSeg {
    u64 pre_off;  // bytes copied verbatim before the payload (block prefix and/or segment header)
    u32 pre_len;
    u32 store;    // 1: store mode, payload is chunked plaintext (compressor.v:297-354)
    u64 in_off, in_len;  // plaintext (store mode)
    u64 pay_off, pay_cap; // coded payload slot in the arena (modeled mode)
    u32 flags;           // bit0: compress() was called (PP byte present)
    u32 last;            // 1: last segment of its block (0xFF follows)
};
struct PackArgs {
    const PackSeg *segs;
    const EncBlock *blocks;
    int n_blocks;
    const u8 *pre;      // all prefix bytes
    const u8 *in;       // plaintext
    const u8 *arena;    // coded payloads
    const u64 *pay_len; // per segment
    const u8 *digests;  // per segment, 20 bytes
    u64 *seg_size;      // scratch: per segment total bytes
    u64 *out_off;       // n_blocks+1
    u8 *out;
    u64 out_cap;        // bytes available at out; blocks that do not fit are not written
};
void launch_pack(const PackArgs &A, int n_segs, cudaStream_t s);

// k_find_blocks: Decompresser.find_block's rolling-hash scan (decompressor.v:227-254)
void launch_find_blocks(const u8 *arc, u64 len, u64 *starts, u32 cap, u32 *count, cudaStream_t s);

// Store-mode blocks (n == 0): decompress_store (decompressor.v:518-587)
void launch_decode_store(const DecodeArgs &A, cudaStream_t s);

// compare stored and computed digests: ok[j] = 1/0, -1 when sha_off == ~0
void launch_sha_compare(const u8 *arc, u64 arc_len, const DecSegRec *recs, int n, const u8 *digests,
                        i32 *ok, cudaStream_t s);

// first `head` bytes of each block [off[b], off[b+1]) gathered for host-side header parsing
void launch_gather_heads(const u8 *arc, const u64 *off, int n, u32 head, u8 *dst, cudaStream_t s);

}  // namespace zg

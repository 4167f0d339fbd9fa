/*
 * zpaqgpu.h -- C ABI of libzpaqgpu, the B200 (sm_100a) ZPAQ block codec.
 *
 * The reference (dy-tea/zpaq-v) has no FFI of its own; its boundary for this path is the V method
 * surface zpaq.Compressor / zpaq.Decompresser.  Each entry point below names the reference
 * interface it stands in for (file:line under /root/reference).  A V host binds these with
 * `#flag -lzpaqgpu`, `#include "zpaqgpu.h"` and `fn C.zpaqgpu_*` declarations.
 *
 * Rules of the boundary: plain pointers and sizes only; the caller owns every host buffer; the
 * library owns device memory and pinned staging; errors are negative status codes, never aborts;
 * there is NO CPU fallback -- without a usable CUDA device every compute call returns
 * ZPAQGPU_E_NODEVICE.  A ctx is bound to one device and is single-threaded.  Several GPUs of one
 * box: either one process (or one ctx) per GPU, or ONE zpaqgpu_multi handle that owns a ctx, a host
 * thread and a stream per device and splits every call into contiguous block ranges (section "several
 * devices" below).  No collective is involved because ZPAQ blocks are independent (compressor.v:84-187
 * re-initialises all model state in start_block).
 */
#ifndef ZPAQGPU_H
#define ZPAQGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

typedef struct zpaqgpu_ctx zpaqgpu_ctx;

enum {
    ZPAQGPU_OK = 0,
    ZPAQGPU_E_NODEVICE = -1, /* no CUDA device / driver; nothing is computed on the host instead */
    ZPAQGPU_E_CUDA = -2,     /* a CUDA call failed; zpaqgpu_last_error() has the text           */
    ZPAQGPU_E_NOSPACE = -3,  /* caller buffer too small; *need holds the required size          */
    ZPAQGPU_E_ARG = -4,      /* bad argument                                                    */
    ZPAQGPU_E_FORMAT = -5,   /* archive bytes the reference decoder would also reject           */
    ZPAQGPU_E_UNSUPPORTED = -6, /* PCOMP/PROG post-processing, sizebits beyond device limits     */
    ZPAQGPU_E_STATE = -7,    /* streaming call in the wrong state (the reference silently returns) */
    ZPAQGPU_E_NOMEM = -8
};

/* Which kernel family a call may use.  AUTO picks the specialised ICM/ISSE-chain kernel when the
 * header has that shape (all of -m1..-m5 do) and the generic all-components kernel otherwise. */
enum { ZPAQGPU_KERNEL_AUTO = 0, ZPAQGPU_KERNEL_GENERIC = 1, ZPAQGPU_KERNEL_CHAIN = 2 };

/* Where the ICM/ISSE hash tables of resident blocks live.  DENSE: the reference's layout, 64<<sizebits
 * bytes per component and block (predictor.v:359-362).  PAGED: a page table per block and 256-byte
 * pages from a pool shared by the wave, mapped on first touch -- same bytes in, same bytes out, but
 * a block costs what it touches (text at -m5: megabytes instead of 2 GiB).  AUTO pages when the
 * dense tables of the batch do not fit in one wave; if the pool runs dry the call falls back to
 * dense waves by itself. */
enum { ZPAQGPU_TABLES_AUTO = 0, ZPAQGPU_TABLES_DENSE = 1, ZPAQGPU_TABLES_PAGED = 2 };

/* ---- lifetime -------------------------------------------------------------------------- */
/* device < 0 selects the current CUDA device.  Replaces Compressor.new()/Decompresser.new()
 * (compressor.v:33, decompressor.v:187) as the owner of all codec state. */
int zpaqgpu_init(zpaqgpu_ctx **out, int device);
void zpaqgpu_destroy(zpaqgpu_ctx *ctx);
const char *zpaqgpu_strerror(int code);
const char *zpaqgpu_last_error(const zpaqgpu_ctx *ctx);
/* Optional tuning: kernel family (above), workspace budget in bytes (0 = 80% of free HBM),
 * CUDA stream to run on (0 = the ctx's own stream). */
int zpaqgpu_set_kernel(zpaqgpu_ctx *ctx, int kernel);
int zpaqgpu_set_table_mode(zpaqgpu_ctx *ctx, int mode);
int zpaqgpu_set_workspace_limit(zpaqgpu_ctx *ctx, uint64_t bytes);
int zpaqgpu_set_stream(zpaqgpu_ctx *ctx, void *cuda_stream);

/* ---- constant data --------------------------------------------------------------------- */
/* get_compression_level(level).hcomp (levels.v:26-375): writes the header bytes
 * "hh hm ph pm n comp.. 0 hcomp.. 0 [0]" and returns their count, or ZPAQGPU_E_NOSPACE. */
int zpaqgpu_level_header(int level, uint8_t *out, int cap);
/* The float-built lookup tables exactly as the device uses them (predictor.v:21-96):
 * squash 4096 x int32, stretch 32768 x int32, state table 1024 x uint8 (statetable.v:15-57). */
int zpaqgpu_tables(int32_t *squash4096, int32_t *stretch32768, uint8_t *state1024);

/* Geometry the library derives from a model header in the levels.v layout, the way
 * Compressor.start_block (compressor.v:96-145) and Predictor.init (predictor.v:292-470) do. */
typedef struct {
    int32_t n, cend, hbegin, hend;   /* components, z.cend, z.hbegin, z.hend                  */
    int32_t hsize;                   /* the two-byte size written after "zPQ" lvl typ         */
    int32_t is_chain;                /* 1: ICM + ISSE chain (+MIX2) shape, specialised kernel  */
    int32_t n_isse, has_mix2;
    int32_t ctx_mode, n_hash;        /* 0 ZPAQL interpreter, 1 level-1 program, 2 hash chain   */
    uint64_t workspace_bytes;        /* HBM bytes of model state per resident block            */
    uint64_t hash_table_bytes;       /* of which ICM/ISSE hash tables                          */
} zpaqgpu_model_info;
int zpaqgpu_describe_model(const uint8_t *header, int header_len, zpaqgpu_model_info *out);

/* ---- batch compression: the fast path -------------------------------------------------- */
/* One ZPAQ block with one segment per input range, exactly the bytes that
 *   Compressor.start_block(level); start_segment(name, comment); for compress(65536) {};
 *   end_segment(); end_block()                       (compressor.v:79-413, cmd/main.v:298-311)
 * writes for that range.  in_off/out_off have n_blocks+1 entries.  names/comments may be NULL
 * (empty strings) or hold NUL-terminated strings per block (entries may be NULL).
 * Returns ZPAQGPU_E_NOSPACE with *out_need set when out_cap is too small. */
int zpaqgpu_compress_blocks(zpaqgpu_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off,
                            int n_blocks, const char *const *names, const char *const *comments,
                            uint8_t *out, uint64_t out_cap, uint64_t *out_off, uint64_t *out_need);
/* Same with caller-supplied model header bytes in the levels.v layout (any header the reference
 * decoder accepts, decompressor.v:278-342: all nine component types and arbitrary HCOMP). */
int zpaqgpu_compress_blocks_header(zpaqgpu_ctx *ctx, const uint8_t *header, int header_len,
                                   const uint8_t *in, const uint64_t *in_off, int n_blocks,
                                   const char *const *names, const char *const *comments,
                                   uint8_t *out, uint64_t out_cap, uint64_t *out_off,
                                   uint64_t *out_need);
/* Device-resident variant: d_in, d_in_off, d_out, d_out_off are device pointers on ctx's device;
 * names/comments are empty.  d_out must hold out_cap bytes; *out_total receives the archive size.
 * Synchronises the stream before returning. */
int zpaqgpu_compress_blocks_dev(zpaqgpu_ctx *ctx, int level, const void *d_in, const void *d_in_off,
                                const uint64_t *h_in_off, int n_blocks, void *d_out,
                                uint64_t out_cap, void *d_out_off, uint64_t *out_total);

/* ---- batch decompression --------------------------------------------------------------- */
/* Decompresser.find_block's locator scan (decompressor.v:227-254) over a whole archive: offsets
 * of the byte AFTER each 16-byte locator+"zPQ" match, ascending.  Returns ZPAQGPU_E_NOSPACE with
 * *n_found = required count when cap is too small. */
int zpaqgpu_find_blocks(zpaqgpu_ctx *ctx, const uint8_t *arc, uint64_t len, uint64_t *starts,
                        int cap, int *n_found);

typedef struct {
    uint64_t block_start;  /* archive offset just after the locator (the level byte)            */
    uint64_t block_end;    /* archive offset just after the block's 0xFF                        */
    uint64_t name_off;     /* archive offset of the NUL-terminated filename                     */
    uint64_t comment_off;  /* archive offset of the NUL-terminated comment                      */
    uint64_t out_off;      /* where the segment's plaintext starts in `out`                     */
    uint64_t out_len;      /* plaintext bytes                                                   */
    int32_t block_index;   /* index of the block among the valid blocks of the archive          */
    int32_t sha1_ok;       /* 1 stored SHA1 matches, 0 mismatch, -1 no checksum (marker 254)    */
} zpaqgpu_segment;

/* for find_block { for find_filename { decompress(-1); read_segment_end } } over the archive
 * (decompressor.v:219-635, cmd/main.v:349-401): every segment's plaintext concatenated in
 * archive order into `out`, one zpaqgpu_segment record each.  size_hints (may be NULL) gives an
 * upper bound of each block's plaintext size to avoid a sizing retry.
 * ZPAQGPU_E_NOSPACE sets *out_need / *n_segs to what is required. */
int zpaqgpu_decompress_archive(zpaqgpu_ctx *ctx, const uint8_t *arc, uint64_t len, uint8_t *out,
                               uint64_t out_cap, uint64_t *out_need, zpaqgpu_segment *segs,
                               int segs_cap, int *n_segs);
/* Device-resident variant for single-segment blocks laid out back to back (what
 * zpaqgpu_compress_blocks_dev produced): block k occupies [h_arc_off[k], h_arc_off[k+1]) of
 * d_arc and its plaintext goes to d_out + h_out_off[k] (capacity h_out_off[k+1]-h_out_off[k]).
 * d_out_len (device, n_blocks x uint64) receives each block's plaintext size; *n_bad counts
 * blocks whose SHA1 or framing failed. */
int zpaqgpu_decompress_blocks_dev(zpaqgpu_ctx *ctx, const void *d_arc, const uint64_t *h_arc_off,
                                  int n_blocks, void *d_out, const uint64_t *h_out_off,
                                  void *d_out_len, int *n_bad);

/* ---- streaming-shaped calls (what the V shim's methods map to) -------------------------- */
/* Compressor.start_block(level) (compressor.v:79).  Data is buffered on the host and the block
 * is coded on the device at block_end; byte output equals the reference's. */
int zpaqgpu_block_begin(zpaqgpu_ctx *ctx, int level);
int zpaqgpu_block_begin_header(zpaqgpu_ctx *ctx, const uint8_t *header, int header_len);
/* Compressor.start_segment(filename, comment) (compressor.v:212) */
int zpaqgpu_segment_begin(zpaqgpu_ctx *ctx, const char *filename, const char *comment);
/* Compressor.compress(n) (compressor.v:259): the shim drains its Reader and passes the bytes.
 * A call with len == 0 still counts as "compress was called" (the PP byte). */
int zpaqgpu_segment_write(zpaqgpu_ctx *ctx, const uint8_t *data, uint64_t len);
/* Compressor.end_segment() (compressor.v:357) */
int zpaqgpu_segment_end(zpaqgpu_ctx *ctx);
/* Compressor.end_block() (compressor.v:402): returns the block's byte count (>=0) written to out,
 * or ZPAQGPU_E_NOSPACE with *need set (call again with a larger buffer; the block is kept). */
int64_t zpaqgpu_block_end(zpaqgpu_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *need);

/* Many blocks one after the other -- the shape of cmd/main.v:288-317, one start_block..end_block per file.
 * A block coded on its own runs ONE chain on a 148-SM GPU; queued blocks are coded together, a chain per
 * block in one launch, and delivered in queue order, so the bytes and their order equal what per-block
 * zpaqgpu_block_end calls give.  The shim's Compressor.end_block() calls zpaqgpu_block_end_queue and, when
 * it returns 1 (queue limits reached: default 1024 blocks or 1 GiB of input, zpaqgpu_stream_batch), or
 * before the output is closed (cmd/main.v:320), zpaqgpu_flush and writes the bytes to its Writer. */
int zpaqgpu_stream_batch(zpaqgpu_ctx *ctx, int max_blocks, uint64_t max_bytes); /* 0 = default */
/* Compressor.end_block() (compressor.v:402), deferred: 0 queued, 1 queued and the queue is full. */
int zpaqgpu_block_end_queue(zpaqgpu_ctx *ctx);
int zpaqgpu_queued(const zpaqgpu_ctx *ctx, int *n_blocks, uint64_t *in_bytes);
/* Codes every queued block; returns the byte count written to out (0 when nothing is queued), or
 * ZPAQGPU_E_NOSPACE with *need set (the coded bytes are kept: call again with a larger buffer). */
int64_t zpaqgpu_flush(zpaqgpu_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *need);

/* ---- jidac front end: fragmentation, fragment hashing, dedup, journaling archive ------------ */
/* Stands in for JidacArchive.create_archive (jidac.v:181-296) and widens it by what the task's
 * north star asks of `jidac add`: content-defined fragmentation with a rolling hash and SHA-1
 * deduplication of fragments, each its own kernel.  The reference always makes ONE fragment per
 * file, never deduplicates and always stores d blocks (jidac.v:94-118); opts {fragment = -1,
 * dedup = 0, block_bytes = 0, level = 0} reproduces exactly those bytes.  The fragmentation rule
 * for fragment >= 0 is upstream zpaq's (order-1 predicted rolling hash, min 64<<fragment, max
 * 8128<<fragment bytes); its source is not part of the reference, parity there is unpinned. */
typedef struct {
    int64_t date;         /* YYYYMMDDHHMMSS as get_jidac_date() makes it (jidac.v:31-35)            */
    int32_t level;        /* method of the d blocks, 0..5; 0 = store, what the reference always uses */
    int32_t fragment;     /* -1: one fragment per file; 0..22: rolling-hash cut, average 1024<<fragment;
                             >22: fixed fragments of 8128<<fragment bytes                           */
    int32_t dedup;        /* 1: fragments with equal SHA-1 and size are stored once                 */
    int32_t reserved;
    uint64_t block_bytes; /* stored fragments are packed into d blocks of up to this many bytes;
                             0 = one d block per stored fragment (the reference: one per file)      */
} zpaqgpu_jidac_opts;

typedef struct {
    uint64_t off, len;    /* byte range inside `in`                                                 */
    uint32_t file;        /* index of the file the fragment belongs to                              */
    uint32_t id;          /* 1-based fragment id as the h and i blocks use it (jidac.v:153-163);
                             duplicates carry the id of their first occurrence                      */
    uint32_t stored;      /* 1: first occurrence, its bytes go into a d block                       */
    uint8_t sha1[20];
} zpaqgpu_fragment;

/* Fragment boundaries, SHA-1 and dedup ids of n_files byte ranges in_off[k]..in_off[k+1] of `in`
 * (files in archive order).  frags receives up to cap records in file order; *n_frags the count
 * (ZPAQGPU_E_NOSPACE when cap is too small, *n_frags then holds the required count), *n_stored how
 * many are first occurrences. */
int zpaqgpu_jidac_fragment(zpaqgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off, int n_files,
                           int fragment, int dedup, zpaqgpu_fragment *frags, int cap, int *n_frags,
                           int *n_stored);

/* A complete journaling archive of the files: c block, d blocks, one h block per d block, i block
 * (jidac.v:181-296; block names jDC<date14><type><num10>, comment "<usize> jDC\x01", jidac.v:47-91).
 * names[k] are NUL-terminated.  ZPAQGPU_E_NOSPACE sets *out_need. */
int zpaqgpu_jidac_add(zpaqgpu_ctx *ctx, const zpaqgpu_jidac_opts *opts, const char *const *names,
                      const uint8_t *in, const uint64_t *in_off, int n_files, uint8_t *out,
                      uint64_t out_cap, uint64_t *out_len, uint64_t *out_need);

/* Restores the files of a journaling archive (north_star: `jidac extract`; the reference has only the
 * writer, jidac.v:181-296, so this follows the layout that writer defines).  Every block is decoded on
 * the device, the h tables (bsize[4] (sha1[20] usize[4])..., jidac.v:229-259) and the i blocks
 * (date[8] name 0 na[4] attr[na] ni[4] ptr[ni][4], jidac.v:262-295; date 0 removes an earlier entry)
 * are followed, the fragments of each file are gathered on the device and copied out once.
 * `out` receives the files back to back in index order, `names` their NUL-terminated names.
 * sha1_ok: 1 when every fragment of the file hashes to the SHA-1 its h table records, else 0.
 * ZPAQGPU_E_NOSPACE sets *out_need / *n_files / *names_need; ZPAQGPU_E_FORMAT: tables inconsistent. */
typedef struct {
    uint64_t name_off;    /* inside `names`                                   */
    uint64_t out_off, out_len;
    int64_t date;         /* YYYYMMDDHHMMSS as stored                         */
    int32_t n_fragments;
    int32_t sha1_ok;
} zpaqgpu_jidac_file;
int zpaqgpu_jidac_extract(zpaqgpu_ctx *ctx, const uint8_t *arc, uint64_t len, uint8_t *out, uint64_t out_cap,
                          uint64_t *out_need, zpaqgpu_jidac_file *files, int files_cap, int *n_files,
                          char *names, uint64_t names_cap, uint64_t *names_need);

typedef struct {
    float h2d_ms, fragment_ms, sha1_ms, dedup_ms, gather_ms; /* front-end stages            */
    float codec_ms;        /* d-block codec kernel                                           */
    float pack_ms, d2h_ms;
    int32_t launches;      /* kernels launched by the last jidac call                        */
    int32_t n_files, n_fragments, n_stored, n_dblocks;
    int32_t reserved;
    uint64_t input_bytes, stored_bytes, archive_bytes;
} zpaqgpu_jidac_stats;
int zpaqgpu_jidac_last_stats(const zpaqgpu_ctx *ctx, zpaqgpu_jidac_stats *out);

/* ---- measurement ------------------------------------------------------------------------ */
typedef struct {
    float init_ms;     /* table zero/fill kernels                      */
    float codec_ms;    /* k_encode_* / k_decode_* (the dominant kernel) */
    float sha1_ms;     /* k_sha1_segments                              */
    float pack_ms;     /* k_scan_sizes + k_pack_blocks / find_blocks   */
    float h2d_ms, d2h_ms;
    int32_t launches;  /* kernels launched by the last call            */
    int32_t codec_launches;
    int32_t waves;     /* table-memory waves the batch was split into  */
    int32_t retries;   /* output-sizing retries                        */
    int32_t kernel;    /* ZPAQGPU_KERNEL_GENERIC or _CHAIN actually used */
    int32_t warps_per_cta;
    uint64_t workspace_bytes_per_block;
    uint64_t pool_bytes_used; /* paged tables: bytes of pool pages mapped by the largest wave */
    int32_t paged;            /* 1 when the last call used paged tables                        */
    int32_t reserved;
} zpaqgpu_stats;
int zpaqgpu_last_stats(const zpaqgpu_ctx *ctx, zpaqgpu_stats *out);

/* ---- several devices of one box behind one handle ---------------------- */
/* One ctx + host thread + stream per device.  A call is split into contiguous ranges balanced by input
 * bytes ("disjoint block ranges to each GPU", no collective), every device stages its range, then every
 * device copies its result to its final place in the caller's buffer: the output is byte for byte what the
 * single-device call writes, in block order.  devices == NULL or n_devices <= 0: all visible devices. */
typedef struct zpaqgpu_multi zpaqgpu_multi;
int zpaqgpu_multi_init(zpaqgpu_multi **out, const int *devices, int n_devices);
void zpaqgpu_multi_destroy(zpaqgpu_multi *m);
int zpaqgpu_multi_device_count(const zpaqgpu_multi *m);
/* the k-th device's context, borrowed (for the zpaqgpu_set_* tuning calls) */
zpaqgpu_ctx *zpaqgpu_multi_ctx(zpaqgpu_multi *m, int k);
const char *zpaqgpu_multi_last_error(const zpaqgpu_multi *m);
/* zpaqgpu_compress_blocks over all devices: same arguments, same bytes. */
int zpaqgpu_multi_compress_blocks(zpaqgpu_multi *m, int level, const uint8_t *in, const uint64_t *in_off,
                                  int n_blocks, const char *const *names, const char *const *comments,
                                  uint8_t *out, uint64_t out_cap, uint64_t *out_off, uint64_t *out_need);
/* zpaqgpu_decompress_archive over all devices: the archive is cut at the first locator at or after
 * len*g/G; when a range does not end cleanly (a block runs across a cut: an archive stored inside an
 * archive; or a damaged block, where the reference's walk stops) one device repeats the whole archive, so
 * the result always equals the single-device call's. */
int zpaqgpu_multi_decompress_archive(zpaqgpu_multi *m, const uint8_t *arc, uint64_t len, uint8_t *out,
                                     uint64_t out_cap, uint64_t *out_need, zpaqgpu_segment *segs,
                                     int segs_cap, int *n_segs);
/* zpaqgpu_jidac_add over all devices -- the one path with an exchange step: every device cuts and hashes a
 * contiguous range of the files, the (SHA-1, length) lists are merged on the host into one fragment table
 * (ids count first occurrences in file order, so a duplicate of a file held by another device is stored once),
 * every device codes the d blocks of the fragments it stores, the first device writes c, h and i blocks.  A d
 * block never spans devices: the block cut depends on the number of devices, the contents of the archive do
 * not; with one device the bytes equal zpaqgpu_jidac_add's. */
int zpaqgpu_multi_jidac_add(zpaqgpu_multi *m, const zpaqgpu_jidac_opts *opts, const char *const *names,
                            const uint8_t *in, const uint64_t *in_off, int n_files, uint8_t *out,
                            uint64_t out_cap, uint64_t *out_len, uint64_t *out_need);
typedef struct {
    int32_t device;          /* CUDA device index                                                */
    int32_t first_unit;      /* first block of the device's range in the last call                */
    int32_t n_units;         /* blocks in the range                                               */
    int32_t fallback_single; /* 1: the last decompress call was repeated on one device            */
    float stage_ms;          /* host time of phase 1 (upload + kernels) on this device            */
    float fetch_ms;          /* host time of phase 2 (copy back)                                  */
    zpaqgpu_stats stats;     /* the device's own counters                                         */
} zpaqgpu_multi_stats;
int zpaqgpu_multi_last_stats(const zpaqgpu_multi *m, int k, zpaqgpu_multi_stats *out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif

/*
 * zpaq_oracle.h -- CPU restatement of the dy-tea/zpaq-v block codec.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline.  libzpaqgpu never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" for compressed bytes at levels 1..5.  The reference is V
 * source, no V toolchain exists in the build container and the reference's own tests hold no
 * golden compressed bytes.  What IS pinned: every exact KAT the reference
 * tests hold for this path (SHA1, state table, oplen, coder init state, level-0 header shape,
 * the "Hello World!" level-1 coder round trip) and the independent survey probe vectors in
 * BASELINE.md section 4.  Every function cites the V file:line it restates.
 */
#ifndef ZPAQ_ORACLE_H
#define ZPAQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tables (predictor.v:21-106, :111-166, statetable.v:15-100, types.v:51-85) ---- */
const int32_t *zo_squash_table(void);   /* 4096 entries, index d+2047 */
const int32_t *zo_stretch_table(void);  /* 32768 entries              */
const int32_t *zo_dt_table(void);       /* 1024 entries               */
const int32_t *zo_dt2k_table(void);     /* 256 entries                */
const uint8_t *zo_state_table(void);    /* 1024 bytes: next0,next1,n0,n1 per state */
int zo_squash(int d);
int zo_stretch(int p);
int zo_st_next(int state, int y);
int zo_st_cminit(int state);
int zo_oplen(int op);
int zo_compsize(int ctype);             /* -1 when ctype is outside the table */
/* levels.v:26-375. Returns the header length (7/26/30/42/57/69) or -1 when cap is too small. */
int zo_level_header(int level, uint8_t *out, int cap);

/* ---- SHA1 (sha1.v:6-146) ---- */
void zo_sha1(const uint8_t *data, size_t n, uint8_t out[20]);

/* ---- growable byte sink standing in for the V Writer ---- */
typedef struct {
    uint8_t *data;
    size_t len, cap;
} zo_buf;
void zo_buf_free(zo_buf *b);

/* ---- Compressor (compressor.v:33-413) ---- */
typedef struct zo_compressor zo_compressor;
zo_compressor *zo_compressor_new(void);
void zo_compressor_free(zo_compressor *c);
/* The V API pulls from a Reader; here input is a borrowed (ptr,len) cursor. */
void zo_compressor_set_input(zo_compressor *c, const uint8_t *data, size_t n);
zo_buf *zo_compressor_output(zo_compressor *c);
void zo_compressor_start_block(zo_compressor *c, int level);
/* Oracle extension (not in the reference API): same as start_block but with caller-supplied
 * header bytes in the get_compression_level() layout "hh hm ph pm n comp.. 0 hcomp.. 0".  It runs
 * the identical parsing code (compressor.v:96-188) so that headers the reference *decoder*
 * accepts (decompressor.v:278-342) can be produced for the generic-component tests. */
void zo_compressor_start_block_header(zo_compressor *c, const uint8_t *hdr, int n);
void zo_compressor_start_segment(zo_compressor *c, const char *filename, const char *comment);
int zo_compressor_compress(zo_compressor *c, int n); /* 1 = n bytes consumed, 0 = EOF/err */
void zo_compressor_end_segment(zo_compressor *c);
void zo_compressor_end_block(zo_compressor *c);

/* ---- Decompresser (decompressor.v:187-640) ---- */
typedef struct zo_decompresser zo_decompresser;
zo_decompresser *zo_decompresser_new(void);
void zo_decompresser_free(zo_decompresser *d);
void zo_decompresser_set_input(zo_decompresser *d, const uint8_t *data, size_t n);
size_t zo_decompresser_input_pos(const zo_decompresser *d);
zo_buf *zo_decompresser_output(zo_decompresser *d);
int zo_decompresser_find_block(zo_decompresser *d);
int zo_decompresser_find_filename(zo_decompresser *d);
const char *zo_decompresser_filename(const zo_decompresser *d);
const char *zo_decompresser_comment(const zo_decompresser *d);
int zo_decompresser_decompress(zo_decompresser *d, int n); /* n<0: all */
void zo_decompresser_read_segment_end(zo_decompresser *d);
/* The reference computes the SHA1 comparison and throws it away (decompressor.v:618-628);
 * the oracle keeps it: 1 match, 0 mismatch, -1 no checksum seen for the last segment. */
int zo_decompresser_last_sha1_ok(const zo_decompresser *d);

/* ---- raw coder + predictor without framing (shape of zpaq_test.v:430-527) ---- */
/* Encodes [optional PP byte 0] data.. EOF flush; returns malloc'd bytes in *out. */
size_t zo_raw_encode(const uint8_t *hdr, int hdr_len, const uint8_t *data, size_t n, int with_pp,
                     uint8_t **out);
/* Decodes until EOF; returns number of bytes written into malloc'd *out. */
size_t zo_raw_decode(const uint8_t *hdr, int hdr_len, const uint8_t *code, size_t n,
                     uint8_t **out);

/* ---- convenience for tests and the CPU baseline ---- */
/* One block, one segment, the way cmd/main.v:288-317 drives the library.  hdr==NULL selects
 * start_block(level).  Returns malloc'd archive bytes. */
size_t zo_compress_block(int level, const uint8_t *hdr, int hdr_len, const uint8_t *data, size_t n,
                         const char *filename, const char *comment, uint8_t **out);
/* Decompress every block/segment of an archive, concatenating the outputs.  *n_bad_sha counts
 * segments whose stored SHA1 did not match.  Returns malloc'd plaintext. */
size_t zo_decompress_archive(const uint8_t *arc, size_t n, uint8_t **out, int *n_segments,
                             int *n_bad_sha);
/* Block-parallel batch over `threads` pthreads (the reference itself is single-threaded and
 * ignores -threads, cmd/main.v:97; this is the harness running independent blocks at once).
 * in_off/out_off have n_blocks+1 entries; out must hold out_cap bytes; returns 0, or -1 when
 * out_cap is too small (then *out_need is the required size). */
int zo_compress_blocks_mt(int level, const uint8_t *in, const uint64_t *in_off, int n_blocks,
                          uint8_t *out, uint64_t out_cap, uint64_t *out_off, uint64_t *out_need,
                          int threads);
int zo_decompress_blocks_mt(const uint8_t *arc, const uint64_t *arc_off, int n_blocks, uint8_t *out,
                            uint64_t out_cap, uint64_t *out_off, uint64_t *out_need, int threads);

#ifdef __cplusplus
}
#endif
#endif

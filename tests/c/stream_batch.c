/* stream_batch.c -- plain-C caller of libzpaqgpu in the shape of cmd/main.v:288-317 (one
 * start_block .. end_block per file).  N files go through
 *   (a) zpaqgpu_compress_blocks                  -- the batch call,
 *   (b) block_begin .. block_end_queue + flush   -- the streaming-shaped calls, queued,
 *   (c) block_begin .. block_end                 -- the streaming-shaped calls, one launch per block (first K files only).
 * The three outputs must be the same bytes.  Prints: n_files input_bytes ms_batch ms_queued equal_ab k ms_single_k equal_ac
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "zpaqgpu.h"

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static uint64_t rng = 0x5A5041512D560009ull;
static uint32_t next_u32(void) {
    rng ^= rng >> 12, rng ^= rng << 25, rng ^= rng >> 27;
    return (uint32_t)((rng * 0x2545F4914F6CDD1Dull) >> 32);
}

#define CHECK(call)                                                                        \
    do {                                                                                   \
        long long rc_ = (long long)(call);                                                 \
        if (rc_ < 0) {                                                                     \
            fprintf(stderr, "%s -> %lld (%s)\n", #call, rc_, zpaqgpu_last_error(ctx));     \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    const int file_kib = argc > 2 ? atoi(argv[2]) : 128;
    const int level = argc > 3 ? atoi(argv[3]) : 2;
    const int k_single = argc > 4 ? atoi(argv[4]) : 4;
    static const char *words[] = {"the", "of", "and", "block", "codec", "zpaq", "stream", "archive", "segment",
                                  "warp", "lane", "table", "hash", "state", "byte", "context"};
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n + 1));
    off[0] = 0;
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + (uint64_t)file_kib * 1024 / 2 + next_u32() % ((uint32_t)file_kib * 1024);
    const uint64_t total = off[n];
    uint8_t *in = (uint8_t *)malloc(total + 64);
    for (uint64_t p = 0; p < total;) {
        const char *w = words[next_u32() & 15];
        const size_t l = strlen(w);
        memcpy(in + p, w, l);
        p += l;
        in[p++] = (next_u32() % 13) ? ' ' : '\n';
    }
    char **names = (char **)malloc(sizeof(char *) * (size_t)n);
    char **comments = (char **)malloc(sizeof(char *) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        names[i] = (char *)malloc(32), comments[i] = (char *)malloc(32);
        snprintf(names[i], 32, "file%05d.txt", i);
        snprintf(comments[i], 32, "%llu bytes", (unsigned long long)(off[i + 1] - off[i]));
    }
    zpaqgpu_ctx *ctx = NULL;
    if (zpaqgpu_init(&ctx, 0) != ZPAQGPU_OK) { fprintf(stderr, "no device\n"); return 3; }
    const uint64_t cap = total + total / 2 + 4096ull * (uint64_t)n;
    uint8_t *a = (uint8_t *)malloc(cap), *b = (uint8_t *)malloc(cap), *c = (uint8_t *)malloc(cap);
    uint64_t *a_off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n + 1));
    uint64_t need = 0;
    /* warm the context (module load, buffer growth) outside the timings */
    CHECK(zpaqgpu_compress_blocks(ctx, level, in, off, n, (const char *const *)names, (const char *const *)comments, a,
                                  cap, a_off, &need));
    double t0 = now_ms();
    CHECK(zpaqgpu_compress_blocks(ctx, level, in, off, n, (const char *const *)names, (const char *const *)comments, a,
                                  cap, a_off, &need));
    const double ms_batch = now_ms() - t0;
    /* (b) queued streaming calls, 64 KiB per compress() call as cmd/main.v:305 does */
    CHECK(zpaqgpu_stream_batch(ctx, n + 1, total + 1));
    t0 = now_ms();
    uint64_t b_len = 0;
    for (int i = 0; i < n; ++i) {
        CHECK(zpaqgpu_block_begin(ctx, level));
        CHECK(zpaqgpu_segment_begin(ctx, names[i], comments[i]));
        uint64_t p = off[i];
        do {
            const uint64_t take = off[i + 1] - p < 65536 ? off[i + 1] - p : 65536;
            CHECK(zpaqgpu_segment_write(ctx, in + p, take));
            p += take;
        } while (p < off[i + 1]);
        CHECK(zpaqgpu_segment_end(ctx));
        CHECK(zpaqgpu_block_end_queue(ctx));
    }
    {
        const long long got = (long long)zpaqgpu_flush(ctx, b, cap, &need);
        if (got < 0) { fprintf(stderr, "flush -> %lld (%s)\n", got, zpaqgpu_last_error(ctx)); return 2; }
        b_len = (uint64_t)got;
    }
    const double ms_queued = now_ms() - t0;
    const int equal_ab = b_len == a_off[n] && memcmp(a, b, b_len) == 0;
    /* (c) one launch per block */
    const int k = k_single < n ? k_single : n;
    t0 = now_ms();
    uint64_t c_len = 0;
    for (int i = 0; i < k; ++i) {
        CHECK(zpaqgpu_block_begin(ctx, level));
        CHECK(zpaqgpu_segment_begin(ctx, names[i], comments[i]));
        CHECK(zpaqgpu_segment_write(ctx, in + off[i], off[i + 1] - off[i]));
        CHECK(zpaqgpu_segment_end(ctx));
        const long long got = (long long)zpaqgpu_block_end(ctx, c + c_len, cap - c_len, &need);
        if (got < 0) { fprintf(stderr, "block_end -> %lld (%s)\n", got, zpaqgpu_last_error(ctx)); return 2; }
        c_len += (uint64_t)got;
    }
    const double ms_single = now_ms() - t0;
    const int equal_ac = c_len == a_off[k] && memcmp(a, c, c_len) == 0;
    printf("%d %llu %.3f %.3f %d %d %.3f %d\n", n, (unsigned long long)total, ms_batch, ms_queued, equal_ab, k, ms_single,
           equal_ac);
    zpaqgpu_destroy(ctx);
    return 0;
}

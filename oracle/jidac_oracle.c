/*
 * jidac_oracle.c -- CPU restatement of the jidac front end: JidacArchive.create_archive of the
 * reference (jidac.v:31-296) and, for the fragmentation + dedup the reference does not have,
 * upstream zpaq's cut rule.
 *
 * TEST INFRASTRUCTURE ONLY (see zpaq_oracle.h): loaded by tests/, smoke() and bench.py's CPU legs.
 *
 * PARITY STATUS
 *  - archive layout (c/d/h/i blocks, names, comments, store mode): pinned by the V source,
 *    jidac.v:47-49, :67-118, :181-296; the reference's tests hold no golden jidac archive.
 *  - fragmentation rule (zo_fragment): "parity unpinned".  It is upstream zpaq 7.15's `add` rule,
 *    whose source is not under /root/reference (the reference cuts ONE fragment per file,
 *    jidac.v:191-200); restated from the published algorithm: per fragment h = c1 = 0, o1[256] = 0,
 *    for each byte c: h = (h + c + 1) * (c == o1[c1] ? 314159265 : 271828182); o1[c1] = c; c1 = c;
 *    cut at EOF, at 8128<<fragment bytes, or when fragment <= 22 and h < 2^(22-fragment) with at
 *    least 64<<fragment bytes.  The GPU kernel and this file are two independent statements of it.
 *  - dedup (fragments with equal SHA-1 and size stored once, ids by first occurrence): the shape
 *    upstream uses; the reference never deduplicates.  Same status.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zpaq_oracle.h"

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

/* Fragment end offsets (relative to data) of one file.  fragment < 0: one fragment, also for an
 * empty file (jidac.v:191-200).  Returns the count; ends may be NULL to count only. */
size_t zo_fragment(const u8 *data, size_t n, int fragment, u64 *ends, size_t cap) {
    size_t cnt = 0;
    if (fragment < 0) {
        if (ends && cap > 0) ends[0] = n;
        return 1;
    }
    const u64 minf = 64ull << fragment, maxf = 8128ull << fragment;
    size_t pos = 0;
    while (pos < n) {
        u8 o1[256];
        memset(o1, 0, sizeof o1);
        u32 h = 0;
        unsigned c1 = 0;
        u64 sz = 0;
        while (pos < n) {
            const unsigned c = data[pos++];
            if (c == o1[c1])
                h = (h + c + 1u) * 314159265u;
            else
                h = (h + c + 1u) * 271828182u;
            o1[c1] = (u8)c;
            c1 = c;
            ++sz;
            if (sz >= maxf || (fragment <= 22 && h < (1u << (22 - fragment)) && sz >= minf)) break;
        }
        if (ends && cnt < cap) ends[cnt] = pos;
        ++cnt;
    }
    return cnt;
}

/* ---- growable bytes ---- */
typedef struct {
    u8 *p;
    size_t len, cap;
} bytes;
static void b_put(bytes *b, const void *src, size_t n) {
    if (b->len + n > b->cap) {
        size_t c = b->cap ? b->cap * 2 : 4096;
        while (c < b->len + n) c *= 2;
        b->p = (u8 *)realloc(b->p, c), b->cap = c;
    }
    if (n) memcpy(b->p + b->len, src, n);
    b->len += n;
}
static void b_le(bytes *b, u64 v, int n) { /* put_u32_le_bytes / put_u64_le_bytes, jidac.v:52-63 */
    u8 t[8];
    for (int i = 0; i < n; i++) t[i] = (u8)(v >> (8 * i));
    b_put(b, t, (size_t)n);
}

/* make_jidac_filename (jidac.v:47-49): "jDC" + itos_pad(date,14) + type + itos_pad(num,10) */
static void jidac_name(char *dst, size_t cap, long long date, char type, u32 num) {
    snprintf(dst, cap, "jDC%014lld%c%010u", date, type, num);
}

/* create_jidac_block / create_data_block (jidac.v:67-118): one block, one segment, comment
 * "<usize> jDC\x01"; the reference always stores (level 0), `level` widens that for d blocks. */
static void put_block(bytes *arc, int level, const u8 *data, size_t n, const char *name) {
    char comment[48];
    snprintf(comment, sizeof comment, "%llu jDC\x01", (unsigned long long)n);
    u8 *blk = NULL;
    const size_t len = zo_compress_block(level, NULL, 0, data, n, name, comment, &blk);
    b_put(arc, blk, len);
    free(blk);
}

typedef struct {
    u64 off, len;
    u32 file, id, stored;
    u8 sha1[20];
} zo_frag;

/* Fragment table of the files in_off[k]..in_off[k+1]; returns malloc'd records, *n their count. */
static zo_frag *frag_table(const u8 *in, const u64 *in_off, int n_files, int fragment, int dedup, size_t *n,
                           u32 *n_stored) {
    size_t cap = 64, cnt = 0;
    zo_frag *fr = (zo_frag *)malloc(cap * sizeof *fr);
    for (int f = 0; f < n_files; f++) {
        const u8 *d = in + in_off[f];
        const size_t len = (size_t)(in_off[f + 1] - in_off[f]);
        const size_t k = zo_fragment(d, len, fragment, NULL, 0);
        u64 *ends = (u64 *)malloc((k + 1) * sizeof *ends);
        zo_fragment(d, len, fragment, ends, k);
        u64 start = 0;
        for (size_t j = 0; j < k; j++) {
            if (cnt == cap) cap *= 2, fr = (zo_frag *)realloc(fr, cap * sizeof *fr);
            zo_frag *x = &fr[cnt++];
            x->off = in_off[f] + start, x->len = ends[j] - start, x->file = (u32)f;
            zo_sha1(in + x->off, (size_t)x->len, x->sha1);
            start = ends[j];
        }
        free(ends);
    }
    /* ids: a.fragments.len after the push (jidac.v:153-163); with dedup an equal earlier fragment
     * lends its id.  Quadratic search is fine for a checker at test sizes; a chained hash keeps the
     * CPU baseline honest at bench sizes. */
    u32 next = 0;
    size_t nb = 64;
    while (nb < 2 * cnt) nb *= 2;
    long *head = (long *)malloc(nb * sizeof *head), *chain = (long *)malloc((cnt + 1) * sizeof *chain);
    for (size_t i = 0; i < nb; i++) head[i] = -1;
    for (size_t i = 0; i < cnt; i++) {
        long rep = -1;
        u32 key;
        memcpy(&key, fr[i].sha1, 4);
        const size_t b = key & (nb - 1);
        if (dedup)
            for (long j = head[b]; j >= 0; j = chain[j])
                if (fr[j].len == fr[i].len && memcmp(fr[j].sha1, fr[i].sha1, 20) == 0) rep = j;
        if (rep >= 0) {
            fr[i].stored = 0, fr[i].id = fr[rep].id;
        } else {
            fr[i].stored = 1, fr[i].id = ++next;
            chain[i] = head[b], head[b] = (long)i;
        }
    }
    free(head), free(chain);
    *n = cnt, *n_stored = next;
    return fr;
}

/* Fragment records as the GPU API reports them (zpaqgpu_fragment has the same layout). */
long zo_jidac_fragment(const u8 *in, const u64 *in_off, int n_files, int fragment, int dedup, void *out,
                       long cap, long *n_stored) {
    size_t n;
    u32 ns;
    zo_frag *fr = frag_table(in, in_off, n_files, fragment, dedup, &n, &ns);
    if (n_stored) *n_stored = (long)ns;
    if ((long)n <= cap && out) memcpy(out, fr, n * sizeof *fr);
    free(fr);
    return (long)n;
}

/* JidacArchive.create_archive (jidac.v:181-296) with the widening described at the top:
 * {fragment = -1, dedup = 0, block_bytes = 0, level = 0} is the reference byte for byte. */
size_t zo_jidac_add(long long date, int level, int fragment, int dedup, u64 block_bytes, const char *const *names,
                    const u8 *in, const u64 *in_off, int n_files, u8 **out) {
    size_t nfr;
    u32 n_stored;
    zo_frag *fr = frag_table(in, in_off, n_files, fragment, dedup, &nfr, &n_stored);
    bytes arc = {0, 0, 0}, dpart = {0, 0, 0}, hpart = {0, 0, 0}, plain = {0, 0, 0}, hc = {0, 0, 0};
    char name[64];
    /* Phase 1 (jidac.v:186-214): d blocks; Phase 4 (:229-259): their h blocks, written later */
    size_t i = 0;
    while (i < nfr) {
        if (!fr[i].stored) {
            i++;
            continue;
        }
        const u32 first_id = fr[i].id;
        plain.len = 0, hc.len = 0;
        b_le(&hc, 0, 4); /* bsize, patched below */
        size_t j = i;
        for (; j < nfr; j++) {
            if (!fr[j].stored) continue;
            if (j > i && (block_bytes == 0 || plain.len + fr[j].len > block_bytes)) break;
            b_put(&plain, in + fr[j].off, (size_t)fr[j].len);
            b_put(&hc, fr[j].sha1, 20);
            b_le(&hc, (u32)fr[j].len, 4);
        }
        const size_t before = dpart.len;
        jidac_name(name, sizeof name, date, 'd', first_id);
        put_block(&dpart, level, plain.p, plain.len, name);
        const u32 csize = (u32)(dpart.len - before);
        for (int k = 0; k < 4; k++) hc.p[k] = (u8)(csize >> (8 * k));
        jidac_name(name, sizeof name, date, 'h', first_id);
        put_block(&hpart, 0, hc.p, hc.len, name);
        i = j;
    }
    /* Phase 2 (:216-225): c block holding the byte count of all d blocks */
    bytes cc = {0, 0, 0};
    b_le(&cc, dpart.len, 8);
    jidac_name(name, sizeof name, date, 'c', n_stored + 1);
    put_block(&arc, 0, cc.p, cc.len, name);
    /* Phase 3 (:226-228), Phase 4 */
    b_put(&arc, dpart.p, dpart.len);
    b_put(&arc, hpart.p, hpart.len);
    /* Phase 5 (:262-295): the i block */
    bytes ic = {0, 0, 0};
    size_t at = 0;
    for (int f = 0; f < n_files; f++) {
        b_le(&ic, (u64)date, 8);
        b_put(&ic, names[f], strlen(names[f]) + 1);
        size_t e = at;
        while (e < nfr && fr[e].file == (u32)f) e++;
        if (date != 0) {
            b_le(&ic, 0, 4);
            b_le(&ic, (u32)(e - at), 4);
            for (size_t k = at; k < e; k++) b_le(&ic, fr[k].id, 4);
        }
        at = e;
    }
    if (ic.len > 0) {
        jidac_name(name, sizeof name, date, 'i', 1);
        put_block(&arc, 0, ic.p, ic.len, name);
    }
    free(fr), free(dpart.p), free(hpart.p), free(plain.p), free(hc.p), free(cc.p), free(ic.p);
    *out = arc.p;
    return arc.len;
}

/*
 * zpaq_oracle.c -- CPU restatement of the dy-tea/zpaq-v block codec (see zpaq_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; "parity unpinned" for compressed bytes (no V toolchain here, no golden
 * archives in the reference).  Build with: gcc -O2 -fwrapv -ffp-contract=off (Makefile).
 *
 * Conventions carried over from V: `int` is 32 bit and wraps (-fwrapv), `>>` on int
 * is arithmetic, int(u32)/u32(int) reinterpret bits, and `& << >> * / %` bind tighter than
 * `+ - | ^`.  Every expression below is parenthesised explicitly.
 */
#include "zpaq_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;
typedef uint16_t u16;

/* ------------------------------------------------------------------------------------------ */
/* byte sink                                                                                  */
/* ------------------------------------------------------------------------------------------ */
static void buf_put(zo_buf *b, int c) {
    if (b->len == b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 4096;
        b->data = (u8 *)realloc(b->data, nc);
        b->cap = nc;
    }
    b->data[b->len++] = (u8)c;
}
void zo_buf_free(zo_buf *b) {
    free(b->data);
    b->data = NULL;
    b->len = b->cap = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* tables                                                                                     */
/* ------------------------------------------------------------------------------------------ */
static i32 g_squash[4096];
static i32 g_stretch[32768];
static i32 g_dt[1024];
static i32 g_dt2k[256];
static u8 g_ns[1024];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* predictor.v:52-70 -- 39-term Taylor series with early exit; not convergent for large |x|. */
static double exp_approx(double x) {
    if (x < -20.0) return 0.0;
    if (x > 20.0) return 485165195.4;
    double result = 1.0, term = 1.0;
    for (int i = 1; i < 40; i++) {
        term *= x / (double)i;
        result += term;
        if (term < 1e-15 && term > -1e-15) break;
    }
    return result;
}

/* predictor.v:169-190 -- atanh series, at most 49 terms. */
static double ln_approx(double x) {
    if (x <= 0.0) return -20.0;
    if (x > 1e9) return 20.0;
    double y = (x - 1.0) / (x + 1.0);
    double y2 = y * y;
    double result = y, term = y;
    for (int i = 1; i < 50; i++) {
        term *= y2;
        result += term / (double)(2 * i + 1);
        if (term < 1e-15 && term > -1e-15) break;
    }
    return 2.0 * result;
}

/* statetable.v:15-57 holds libzpaq's 1024-byte `sns` table as literal data.  The table is the
 * output of libzpaq's published StateTable generator; it is regenerated here and the tests check
 * it byte for byte against the reference's literal (tests/test_oracle_reference_data.py). */
static int st_num_states(int n0, int n1) {
    static const int bound[6] = {20, 48, 15, 8, 6, 5};
    if (n0 < n1) return st_num_states(n1, n0);
    if (n0 < 0 || n1 < 0 || n1 >= 6 || n0 > bound[n1]) return 0;
    return 1 + (n1 > 0 && n0 + n1 <= 17);
}
static void st_discount(int *n0) {
    *n0 = (*n0 >= 1) + (*n0 >= 2) + (*n0 >= 3) + (*n0 >= 4) + (*n0 >= 5) + (*n0 >= 7) + (*n0 >= 8);
}
static void st_next_state(int *n0, int *n1, int y) {
    if (*n0 < *n1) {
        st_next_state(n1, n0, 1 - y);
        return;
    }
    if (y) {
        ++*n1;
        st_discount(n0);
    } else {
        ++*n0;
        st_discount(n1);
    }
    while (!st_num_states(*n0, *n1)) {
        if (*n1 < 2)
            --*n0;
        else {
            *n0 = (*n0 * (*n1 - 1) + (*n1 / 2)) / *n1;
            --*n1;
        }
    }
}
static void build_state_table(void) {
    enum { N = 50 };
    static u8 t[N][N][2];
    int state = 0;
    memset(t, 0, sizeof(t));
    for (int i = 0; i < N; ++i)
        for (int n1 = 0; n1 <= i; ++n1) {
            int n0 = i - n1;
            int n = st_num_states(n0, n1);
            if (n) {
                t[n0][n1][0] = (u8)state;
                t[n0][n1][1] = (u8)(state + n - 1);
                state += n;
            }
        }
    memset(g_ns, 0, sizeof(g_ns));
    for (int n0 = 0; n0 < N; ++n0)
        for (int n1 = 0; n1 < N; ++n1)
            for (int y = 0; y < st_num_states(n0, n1); ++y) {
                int s = t[n0][n1][y];
                int s0 = n0, s1 = n1;
                st_next_state(&s0, &s1, 0);
                g_ns[s * 4 + 0] = t[s0][s1][0];
                s0 = n0, s1 = n1;
                st_next_state(&s0, &s1, 1);
                g_ns[s * 4 + 1] = t[s0][s1][1];
                g_ns[s * 4 + 2] = (u8)n0;
                g_ns[s * 4 + 3] = (u8)n1;
            }
}

static void init_tables(void) {
    /* predictor.v:21-49 */
    memset(g_squash, 0, sizeof(g_squash));
    for (int i = -2047; i <= 2047; i++) {
        double d = (double)i / 64.0;
        if (d < -20.0) d = -20.0;
        if (d > 20.0) d = 20.0;
        double e;
        if (d >= 0) {
            e = 1.0 / (1.0 + exp_approx(-d));
        } else {
            double tmp = exp_approx(d);
            e = tmp / (1.0 + tmp);
        }
        i32 v = (i32)(32767.0 * e + 0.5);
        g_squash[i + 2047] = v < 1 ? 1 : (v > 32767 ? 32767 : v);
    }
    /* predictor.v:73-96 */
    for (int i = 0; i < 32768; i++) {
        double p = (double)i / 32767.0;
        if (p <= 0.0)
            g_stretch[i] = -2047;
        else if (p >= 1.0)
            g_stretch[i] = 2047;
        else {
            double ln_odds = ln_approx(p / (1.0 - p));
            i32 v = (i32)(ln_odds * 64.0);
            g_stretch[i] = v < -2047 ? -2047 : (v > 2047 ? 2047 : v);
        }
    }
    /* predictor.v:99-106 */
    for (int i = 0; i < 256; i++) g_dt2k[i] = 2048 - 2048 / (i + 1);
    /* predictor.v:111-166 is the literal list of (1<<17)/(i*2+3)*2 (comment at :109). */
    for (int i = 0; i < 1024; i++) g_dt[i] = (1 << 17) / (i * 2 + 3) * 2;
    build_state_table();
}
static void ensure_tables(void) { pthread_once(&g_once, init_tables); }

const i32 *zo_squash_table(void) { ensure_tables(); return g_squash; }
const i32 *zo_stretch_table(void) { ensure_tables(); return g_stretch; }
const i32 *zo_dt_table(void) { ensure_tables(); return g_dt; }
const i32 *zo_dt2k_table(void) { ensure_tables(); return g_dt2k; }
const u8 *zo_state_table(void) { ensure_tables(); return g_ns; }

/* predictor.v:193-202 (index clamps to [0,4093], Q1) */
static inline i32 squash(i32 d) {
    i32 idx = d + 2047;
    if (idx < 0) idx = 0;
    if (idx >= 4094) idx = 4093;
    return g_squash[idx];
}
/* predictor.v:205-214 */
static inline i32 stretch(i32 p) {
    i32 idx = p;
    if (idx < 1) idx = 1;
    if (idx >= 32768) idx = 32767;
    return g_stretch[idx];
}
/* predictor.v:217-236 */
static inline i32 clamp2k(i32 x) { return x < -2048 ? -2048 : (x > 2047 ? 2047 : x); }
static inline i32 clamp512k(i32 x) { return x < -262144 ? -262144 : (x > 262143 ? 262143 : x); }
/* statetable.v:75-100 */
static inline i32 st_next(i32 state, i32 y) {
    if (state < 0 || state >= 256) return 0;
    i32 idx = state * 4 + y;
    if (idx < 0 || idx >= 1024) return 0;
    return g_ns[idx];
}
static inline i32 st_cminit(i32 state) {
    if (state < 0 || state >= 256) return 1 << 22;
    u32 n0 = g_ns[state * 4 + 2], n1 = g_ns[state * 4 + 3];
    return (i32)(((n1 * 2 + 1) << 22) / (n0 + n1 + 1));
}
int zo_squash(int d) { ensure_tables(); return squash(d); }
int zo_stretch(int p) { ensure_tables(); return stretch(p); }
int zo_st_next(int s, int y) { ensure_tables(); return st_next(s, y); }
int zo_st_cminit(int s) { ensure_tables(); return st_cminit(s); }

/* types.v:51-85 */
static const int k_compsize[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};
int zo_compsize(int t) { return (t < 0 || t >= 10) ? -1 : k_compsize[t]; }
static inline int oplen(u8 op) { return op == 255 ? 3 : ((op & 7) == 7 ? 2 : 1); }
int zo_oplen(int op) { return oplen((u8)op); }

/* levels.v:40-375.  The five model headers, rebuilt from their parameters; the tests compare the
 * bytes against the reference's literals. */
int zo_level_header(int level, u8 *out, int cap) {
    u8 h[80];
    int n = 0;
    if (level == 0) {
        memset(h, 0, 7);
        n = 7;
    } else if (level < 1 || level > 5 || level == 1) {
        /* levels.v:34 -- any other level falls back to level 1 */
        static const u8 m1[26] = {1,  2,  0,  0,  2,   3,  16, 8,  19, 0,  0,   96, 4,
                                  28, 59, 10, 59, 112, 25, 10, 59, 10, 59, 112, 56, 0};
        memcpy(h, m1, 26);
        n = 26;
    } else {
        static const u8 hh[6] = {0, 0, 9, 10, 12, 14}, hm[6] = {0, 0, 16, 18, 20, 22};
        static const u8 bits[6] = {0, 0, 16, 18, 20, 22}, nisse[6] = {0, 0, 2, 4, 5, 7};
        int mix2 = level >= 4;
        int ncomp = 1 + nisse[level] + mix2;
        h[n++] = hh[level], h[n++] = hm[level], h[n++] = 0, h[n++] = 0, h[n++] = (u8)ncomp;
        h[n++] = 3, h[n++] = bits[level];
        for (int i = 0; i < nisse[level]; i++) h[n++] = 8, h[n++] = bits[level], h[n++] = (u8)i;
        if (mix2) {
            h[n++] = 6, h[n++] = level == 4 ? 16 : 18;
            h[n++] = (u8)(nisse[level] - 1), h[n++] = nisse[level], h[n++] = 24, h[n++] = 255;
        }
        h[n++] = 0;
        h[n++] = 74, h[n++] = 18, h[n++] = 104, h[n++] = 95, h[n++] = 0;
        for (int i = 0; i < ncomp; i++) {
            h[n++] = 59, h[n++] = 112;
            if (i + 1 < ncomp) h[n++] = 25;
        }
        h[n++] = 56, h[n++] = 0, h[n++] = 0;
    }
    if (n > cap) return -1;
    memcpy(out, h, (size_t)n);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* SHA1 (sha1.v:6-146)                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    u64 len0;
    u32 h[5];
    u8 buf[64];
    int bufn;
    int final;
} sha1_t;
static inline u32 rotl(u32 x, u32 n) { return (x << n) | (x >> (32 - n)); }
static void sha1_init(sha1_t *s) {
    s->len0 = 0, s->bufn = 0, s->final = 0;
    s->h[0] = 0x67452301u, s->h[1] = 0xEFCDAB89u, s->h[2] = 0x98BADCFEu, s->h[3] = 0x10325476u,
    s->h[4] = 0xC3D2E1F0u;
}
static void sha1_block(sha1_t *s) {
    u32 w[80];
    for (int i = 0; i < 16; i++)
        w[i] = ((u32)s->buf[i * 4] << 24) | ((u32)s->buf[i * 4 + 1] << 16) |
               ((u32)s->buf[i * 4 + 2] << 8) | (u32)s->buf[i * 4 + 3];
    for (int i = 16; i < 80; i++) w[i] = rotl(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    u32 a = s->h[0], b = s->h[1], c = s->h[2], d = s->h[3], e = s->h[4];
    for (int i = 0; i < 80; i++) {
        u32 f, k;
        if (i < 20)
            f = (b & c) | ((~b) & d), k = 0x5A827999u;
        else if (i < 40)
            f = b ^ c ^ d, k = 0x6ED9EBA1u;
        else if (i < 60)
            f = (b & c) | (b & d) | (c & d), k = 0x8F1BBCDCu;
        else
            f = b ^ c ^ d, k = 0xCA62C1D6u;
        u32 temp = rotl(a, 5) + f + e + k + w[i];
        e = d, d = c, c = rotl(b, 30), b = a, a = temp;
    }
    s->h[0] += a, s->h[1] += b, s->h[2] += c, s->h[3] += d, s->h[4] += e;
}
static inline void sha1_put(sha1_t *s, int c) {
    if (s->final) return;
    s->buf[s->bufn++] = (u8)c;
    s->len0 += 8;
    if (s->bufn == 64) sha1_block(s), s->bufn = 0;
}
static void sha1_result(sha1_t *s, u8 out[20]) {
    if (!s->final) {
        s->buf[s->bufn++] = 0x80;
        if (s->bufn > 56) {
            while (s->bufn < 64) s->buf[s->bufn++] = 0;
            sha1_block(s);
            s->bufn = 0;
        }
        while (s->bufn < 56) s->buf[s->bufn++] = 0;
        for (int i = 7; i >= 0; i--) s->buf[s->bufn++] = (u8)(s->len0 >> (i * 8));
        sha1_block(s);
        s->final = 1;
    }
    for (int i = 0; i < 5; i++) {
        out[i * 4] = (u8)(s->h[i] >> 24), out[i * 4 + 1] = (u8)(s->h[i] >> 16);
        out[i * 4 + 2] = (u8)(s->h[i] >> 8), out[i * 4 + 3] = (u8)s->h[i];
    }
}
void zo_sha1(const u8 *data, size_t n, u8 out[20]) {
    sha1_t s;
    sha1_init(&s);
    for (size_t i = 0; i < n; i++) sha1_put(&s, data[i]);
    sha1_result(&s, out);
}

/* ------------------------------------------------------------------------------------------ */
/* ZPAQL VM (zpaql.v)                                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    u32 a, b, c, d;
    i32 f, pc;
    u8 *m;
    u32 mlen;
    u32 *h;
    u32 hlen;
    u32 r[256];
    u8 *header;
    i32 hlen_hdr; /* header.len */
    i32 cend, hbegin, hend;
} zvm;

static void vm_new(zvm *z) { memset(z, 0, sizeof(*z)); }
static void vm_free(zvm *z) {
    free(z->m), free(z->h), free(z->header);
    memset(z, 0, sizeof(*z));
}
/* zpaql.v:54-70 */
static void vm_clear(zvm *z) {
    z->a = z->b = z->c = z->d = 0;
    z->f = 0, z->pc = 0;
    if (z->m) memset(z->m, 0, z->mlen);
    if (z->h) memset(z->h, 0, (size_t)z->hlen * 4);
    memset(z->r, 0, sizeof(z->r));
}
static void vm_set_header(zvm *z, const u8 *hdr, int n) {
    free(z->header);
    z->header = (u8 *)malloc((size_t)(n > 0 ? n : 1));
    if (n > 0) memcpy(z->header, hdr, (size_t)n);
    z->hlen_hdr = n;
}
/* zpaql.v:74-83: H is (re)allocated only when 0 < hh < 32 */
static void vm_inith(zvm *z) {
    if (z->hlen_hdr < 2) return;
    int hh = z->header[0];
    if (hh > 0 && hh < 32) {
        free(z->h);
        z->hlen = 1u << hh;
        z->h = (u32 *)calloc(z->hlen, 4);
    }
}
/* zpaql.v:86-96 */
static void vm_initp(zvm *z) {
    if (z->hlen_hdr < 2) return;
    int hm = z->header[1];
    if (hm > 0 && hm < 32) {
        free(z->m);
        z->mlen = 1u << hm;
        z->m = (u8 *)calloc(z->mlen, 1);
    }
    z->pc = z->hbegin;
}
/* zpaql.v:178-211 */
static inline u8 m_get(const zvm *z, u32 i) { return z->mlen ? z->m[i & (z->mlen - 1)] : 0; }
static inline void m_set(zvm *z, u32 i, u8 v) {
    if (z->mlen) z->m[i & (z->mlen - 1)] = v;
}
static inline u32 h_get(const zvm *z, u32 i) { return z->hlen ? z->h[i & (z->hlen - 1)] : 0; }
static inline void h_set(zvm *z, u32 i, u32 v) {
    if (z->hlen) z->h[i & (z->hlen - 1)] = v;
}

/* zpaql.v:215-954.  Returns 0 to stop the run. */
static int vm_execute(zvm *z) {
    if (z->pc < z->hbegin || z->pc >= z->hend) return 0;
    u8 op = z->header[z->pc];
    z->pc++;
    i32 operand = 0;
    if (oplen(op) == 2 && z->pc < z->hlen_hdr) {
        operand = z->header[z->pc];
        z->pc++;
    } else if (oplen(op) == 3 && z->pc + 1 < z->hlen_hdr) {
        operand = (i32)z->header[z->pc] + (i32)z->header[z->pc + 1] * 256;
        z->pc += 2;
    }
    if (op >= 64 && op < 120) { /* X=Y, op = 64 + 8*X + Y (zpaql.v:365-532) */
        int y = op & 7, x = (op >> 3) & 7;
        u32 v;
        switch (y) {
        case 0: v = z->a; break;
        case 1: v = z->b; break;
        case 2: v = z->c; break;
        case 3: v = z->d; break;
        case 4: v = m_get(z, z->b); break;
        case 5: v = m_get(z, z->c); break;
        case 6: v = h_get(z, z->d); break;
        default: v = (u32)operand; break;
        }
        switch (x) {
        case 0: z->a = v; break;
        case 1: z->b = v; break;
        case 2: z->c = v; break;
        case 3: z->d = v; break;
        case 4: m_set(z, z->b, (u8)v); break;
        case 5: m_set(z, z->c, (u8)v); break;
        default: h_set(z, z->d, v); break;
        }
        return 1;
    }
    if (op >= 128 && op < 240) { /* a op= Y (zpaql.v:534-941) */
        int y = op & 7, g = (op - 128) >> 3;
        u32 v;
        switch (y) {
        case 0: v = z->a; break;
        case 1: v = z->b; break;
        case 2: v = z->c; break;
        case 3: v = z->d; break;
        case 4: v = m_get(z, z->b); break;
        case 5: v = m_get(z, z->c); break;
        case 6: v = h_get(z, z->d); break;
        default: v = (u32)operand; break;
        }
        switch (g) {
        case 0: z->a += v; break;
        case 1: z->a -= v; break;
        case 2: z->a *= v; break;
        case 3: if (v != 0) z->a /= v; break; /* divide by zero leaves a unchanged */
        case 4: if (v != 0) z->a %= v; break;
        case 5: z->a &= v; break;
        case 6: z->a &= ~v; break;
        case 7: z->a |= v; break;
        case 8: z->a ^= v; break;
        case 9: z->a <<= (v & 31); break;
        case 10: z->a >>= (v & 31); break;
        case 11: z->f = (z->a == v); break;
        case 12: z->f = (z->a < v); break;
        default: z->f = (z->a > v); break;
        }
        return 1;
    }
    u32 t;
    switch (op) {
    case 0: break;
    case 1: z->a++; break;
    case 2: z->a--; break;
    case 3: z->a = ~z->a; break;
    case 4: z->a = 0; break;
    case 7: z->a = z->r[operand & 255]; break;
    case 8: t = z->a, z->a = z->b, z->b = t; break;
    case 9: z->b++; break;
    case 10: z->b--; break;
    case 11: z->b = ~z->b; break;
    case 12: z->b = 0; break;
    case 15: z->b = z->r[operand & 255]; break;
    case 16: t = z->a, z->a = z->c, z->c = t; break;
    case 17: z->c++; break;
    case 18: z->c--; break;
    case 19: z->c = ~z->c; break;
    case 20: z->c = 0; break;
    case 23: z->c = z->r[operand & 255]; break;
    case 24: t = z->a, z->a = z->d, z->d = t; break;
    case 25: z->d++; break;
    case 26: z->d--; break;
    case 27: z->d = ~z->d; break;
    case 28: z->d = 0; break;
    case 31: z->d = z->r[operand & 255]; break;
    case 32: t = m_get(z, z->b), m_set(z, z->b, (u8)z->a), z->a = t; break;
    case 33: m_set(z, z->b, (u8)(m_get(z, z->b) + 1)); break;
    case 34: m_set(z, z->b, (u8)(m_get(z, z->b) - 1)); break;
    case 35: m_set(z, z->b, (u8)~m_get(z, z->b)); break;
    case 36: m_set(z, z->b, 0); break;
    case 39: if (z->f != 0) z->pc += ((operand + 128) & 255) - 127; break; /* Q5 */
    case 40: t = m_get(z, z->c), m_set(z, z->c, (u8)z->a), z->a = t; break;
    case 41: m_set(z, z->c, (u8)(m_get(z, z->c) + 1)); break;
    case 42: m_set(z, z->c, (u8)(m_get(z, z->c) - 1)); break;
    case 43: m_set(z, z->c, (u8)~m_get(z, z->c)); break;
    case 44: m_set(z, z->c, 0); break;
    case 47: if (z->f == 0) z->pc += ((operand + 128) & 255) - 127; break;
    case 48: t = h_get(z, z->d), h_set(z, z->d, z->a), z->a = t; break;
    case 49: h_set(z, z->d, h_get(z, z->d) + 1); break;
    case 50: h_set(z, z->d, h_get(z, z->d) - 1); break;
    case 51: h_set(z, z->d, ~h_get(z, z->d)); break;
    case 52: h_set(z, z->d, 0); break;
    case 55: z->r[operand & 255] = z->a; break;
    case 56: return 0; /* HALT */
    case 57: break;    /* OUT: appends to an outbuf nobody reads on the HCOMP path (zpaql.v:149-157) */
    case 59: z->a = (z->a + (u32)m_get(z, z->b) + 512) * 773; break;
    case 60: h_set(z, z->d, (h_get(z, z->d) + z->a + 512) * 773); break;
    case 63: z->pc += ((operand + 128) & 255) - 127; break;
    case 255: /* zpaql.v:942-947: reads the two bytes before pc whether or not pc was advanced */
        z->pc = z->hbegin + (i32)z->header[z->pc - 2] + (i32)z->header[z->pc - 1] * 256;
        if (z->pc >= z->hend) return 0;
        break;
    default: return 0;
    }
    return 1;
}
/* zpaql.v:167-175 */
static void vm_run(zvm *z, u32 input) {
    z->a = input;
    z->pc = z->hbegin;
    while (z->pc < z->hend && z->pc >= z->hbegin)
        if (!vm_execute(z)) break;
}

/* Header geometry as Compressor.start_block computes it (compressor.v:96-145). */
static void vm_parse_geometry(zvm *z) {
    int len = z->hlen_hdr;
    if (len >= 5) {
        int n = z->header[4];
        int pos = 5;
        for (int i = 0; i < n && pos < len; i++) {
            int ctype = z->header[pos];
            if (ctype < 0 || ctype >= 10) break;
            pos += k_compsize[ctype];
        }
        z->cend = pos;
        if (pos < len && z->header[pos] == 0) pos++;
        z->hbegin = pos;
        while (pos < len) {
            u8 op = z->header[pos];
            if (op == 0) break;
            pos++;
            if ((op & 7) == 7) pos += (op == 63) ? 2 : 1; /* Q14 */
        }
        z->hend = pos;
    } else {
        z->cend = z->hbegin = z->hend = len;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Predictor (predictor.v:239-833)                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    i32 ctype;
    u32 *cm;
    i32 cm_len;
    u8 *ht;
    i32 ht_len;
    u16 *a16;
    i32 a16_len;
    i32 a, b, c;
    u32 cxt;
    i32 limit;
} comp_t;

typedef struct {
    u32 c8, hmap4;
    u32 *h;
    i32 *p;
    comp_t *comp;
    i32 n;
    zvm *z;
} pred_t;

static void pred_free(pred_t *pr) {
    for (int i = 0; i < pr->n; i++) free(pr->comp[i].cm), free(pr->comp[i].ht), free(pr->comp[i].a16);
    free(pr->comp), free(pr->h), free(pr->p);
    memset(pr, 0, sizeof(*pr));
}

/* predictor.v:292-470 */
static void pred_init(pred_t *pr, zvm *z) {
    ensure_tables();
    pred_free(pr);
    pr->z = z;
    pr->c8 = 1, pr->hmap4 = 1;
    if (z->hlen_hdr < 5) return;
    int n = z->header[4];
    if (n == 0) return;
    pr->n = n;
    pr->comp = (comp_t *)calloc((size_t)n, sizeof(comp_t));
    pr->p = (i32 *)calloc((size_t)n, sizeof(i32));
    pr->h = (u32 *)calloc((size_t)n, sizeof(u32));
    const u8 *hd = z->header;
    int cp = 5;
    for (int i = 0; i < n && cp < z->cend; i++) {
        comp_t *cr = &pr->comp[i];
        int ctype = hd[cp];
        cr->ctype = ctype;
        switch (ctype) {
        case 1:
            cr->a = hd[cp + 1];
            cp += 2;
            break;
        case 2: {
            cr->a = hd[cp + 1];
            cr->limit = (i32)hd[cp + 2] * 4;
            i32 size = 1 << cr->a;
            cr->cm = (u32 *)malloc((size_t)size * 4), cr->cm_len = size;
            for (i32 j = 0; j < size; j++) cr->cm[j] = 0x80000000u;
            cp += 3;
            break;
        }
        case 3: {
            cr->a = hd[cp + 1];
            i32 size = 16 << (cr->a + 2);
            cr->cm = (u32 *)malloc(256 * 4), cr->cm_len = 256;
            cr->ht = (u8 *)calloc((size_t)size, 1), cr->ht_len = size;
            for (int j = 0; j < 256; j++) cr->cm[j] = (u32)st_cminit(j);
            cp += 2;
            break;
        }
        case 4:
            cr->a = hd[cp + 1];
            cr->b = hd[cp + 2];
            cr->cm_len = 1 << cr->a, cr->cm = (u32 *)calloc((size_t)cr->cm_len, 4);
            cr->ht_len = 1 << cr->b, cr->ht = (u8 *)calloc((size_t)cr->ht_len, 1);
            cr->limit = 0, cr->c = 0, cr->cxt = 0;
            cp += 3;
            break;
        case 5:
            cr->a = hd[cp + 1], cr->b = hd[cp + 2], cr->c = hd[cp + 3];
            cp += 4;
            break;
        case 6: {
            cr->a = hd[cp + 1];
            i32 size = 1 << cr->a;
            cr->b = hd[cp + 2];
            cr->c = size;
            cr->a16 = (u16 *)malloc((size_t)size * 2), cr->a16_len = size;
            for (i32 j = 0; j < size; j++) cr->a16[j] = 32768;
            cr->cm = (u32 *)malloc(16), cr->cm_len = 4;
            cr->cm[0] = hd[cp + 2], cr->cm[1] = hd[cp + 3], cr->cm[2] = hd[cp + 4], cr->cm[3] = hd[cp + 5];
            cp += 6;
            break;
        }
        case 7: {
            cr->a = hd[cp + 1];
            i32 size = 1 << cr->a;
            i32 j = hd[cp + 2], m = hd[cp + 3], rate = hd[cp + 4], mask = hd[cp + 5];
            cr->b = j, cr->c = size, cr->limit = m;
            cr->ht = (u8 *)malloc(2), cr->ht_len = 2;
            cr->ht[0] = (u8)rate, cr->ht[1] = (u8)mask;
            cr->cm_len = size * m;
            cr->cm = (u32 *)malloc((size_t)(cr->cm_len > 0 ? cr->cm_len : 1) * 4);
            /* m == 0 would divide by zero in the reference (predictor.v:426); callers avoid it */
            for (i32 k = 0; k < size * m; k++) cr->cm[k] = (u32)(65536 / m) << 8;
            cp += 6;
            break;
        }
        case 8: {
            cr->a = hd[cp + 1], cr->b = hd[cp + 2];
            i32 size = 16 << (cr->a + 2);
            cr->ht = (u8 *)calloc((size_t)size, 1), cr->ht_len = size;
            cr->cm = (u32 *)malloc(512 * 4), cr->cm_len = 512;
            for (int k = 0; k < 256; k++) {
                cr->cm[k * 2] = (u32)(1 << 15);
                cr->cm[k * 2 + 1] = (u32)clamp512k(stretch(st_cminit(k) >> 8) * 1024);
            }
            cp += 3;
            break;
        }
        case 9: {
            cr->a = hd[cp + 1], cr->b = hd[cp + 2];
            i32 size = 1 << cr->a;
            cr->cm_len = size * 32, cr->cm = (u32 *)malloc((size_t)cr->cm_len * 4);
            cr->limit = (i32)hd[cp + 4] * 4;
            i32 start = hd[cp + 3];
            for (i32 k = 0; k < size * 32; k++) {
                i32 q = (k & 31) * 64 - 992;
                cr->cm[k] = ((u32)squash(q) << 17) | (u32)start;
            }
            cp += 5;
            break;
        }
        default:
            cp += 1;
            break;
        }
    }
}

/* predictor.v:827-833 */
static void pred_reset(pred_t *pr) {
    pr->c8 = 1, pr->hmap4 = 1;
    for (int i = 0; i < pr->n; i++) pr->h[i] = 0;
}

/* predictor.v:495-532 */
static i32 find_ht(u8 *ht, i32 ht_len, i32 sizebits, u32 cxt) {
    i32 chk = (i32)((cxt >> sizebits) & 255);
    i32 h0 = (i32)((cxt * 16) & (u32)(ht_len - 16));
    if (ht[h0] == (u8)chk) return h0;
    i32 h1 = h0 ^ 16;
    if (ht[h1] == (u8)chk) return h1;
    i32 h2 = h0 ^ 32;
    if (ht[h2] == (u8)chk) return h2;
    if (ht[h0 + 1] <= ht[h1 + 1] && ht[h0 + 1] <= ht[h2 + 1]) {
        memset(ht + h0, 0, 16), ht[h0] = (u8)chk;
        return h0;
    } else if (ht[h1 + 1] < ht[h2 + 1]) {
        memset(ht + h1, 0, 16), ht[h1] = (u8)chk;
        return h1;
    } else {
        memset(ht + h2, 0, 16), ht[h2] = (u8)chk;
        return h2;
    }
}

/* predictor.v:536-668 */
static i32 pred_predict(pred_t *pr) {
    i32 n = pr->n;
    if (n == 0) return 16384;
    i32 *p = pr->p;
    for (i32 i = 0; i < n; i++) {
        comp_t *cr = &pr->comp[i];
        switch (cr->ctype) {
        case 1:
            p[i] = (cr->a - 128) * 16;
            break;
        case 2: {
            cr->cxt = pr->h[i] ^ pr->hmap4;
            i32 idx = (i32)cr->cxt & (cr->cm_len - 1);
            p[i] = stretch((i32)(cr->cm[idx] >> 17));
            break;
        }
        case 3:
            if (pr->c8 == 1 || (pr->c8 & 0xf0) == 16)
                cr->c = find_ht(cr->ht, cr->ht_len, cr->a + 2, pr->h[i] + 16 * pr->c8);
            cr->cxt = cr->ht[cr->c + (i32)(pr->hmap4 & 15)];
            p[i] = stretch((i32)(cr->cm[cr->cxt] >> 8));
            break;
        case 4:
            if (cr->a == 0) {
                p[i] = 0;
            } else {
                i32 idx = (cr->limit - cr->b) & (cr->ht_len - 1);
                cr->c = (i32)((cr->ht[idx] >> (7 - (i32)cr->cxt)) & 1);
                i32 weight = g_dt2k[cr->a & 255];
                p[i] = stretch((weight * (cr->c * -2 + 1)) & 32767);
            }
            break;
        case 5: {
            i32 j = cr->a, k = cr->b, wt = cr->c;
            p[i] = (j < n && k < n) ? ((p[j] * wt + p[k] * (256 - wt)) >> 8) : 0;
            break;
        }
        case 6: {
            i32 j = (i32)cr->cm[0], k = (i32)cr->cm[1], mask = (i32)cr->cm[3];
            cr->cxt = (pr->h[i] + (pr->c8 & (u32)mask)) & (u32)(cr->c - 1);
            i32 w = cr->a16[cr->cxt];
            p[i] = (j < n && k < n) ? clamp2k((w * p[j] + (65536 - w) * p[k]) >> 16) : 0;
            break;
        }
        case 7: {
            i32 j = cr->b, m = cr->limit, mask = cr->ht[1];
            cr->cxt = (u32)(((i32)pr->h[i] + ((i32)pr->c8 & mask)) & (cr->c - 1));
            i32 idx = (i32)cr->cxt * m;
            i32 sum = 0;
            for (i32 l = 0; l < m && (j + l) < n; l++) {
                i32 wt = (i32)cr->cm[idx + l] >> 8;
                sum += wt * p[j + l];
            }
            p[i] = clamp2k(sum >> 8);
            break;
        }
        case 8: {
            if (pr->c8 == 1 || (pr->c8 & 0xf0) == 16)
                cr->c = find_ht(cr->ht, cr->ht_len, cr->a + 2, pr->h[i] + 16 * pr->c8);
            cr->cxt = cr->ht[cr->c + (i32)(pr->hmap4 & 15)];
            i32 wt0 = (i32)cr->cm[cr->cxt * 2], wt1 = (i32)cr->cm[cr->cxt * 2 + 1];
            i32 j = cr->b;
            p[i] = (j < n) ? clamp2k((wt0 * p[j] + wt1 * 64) >> 16) : clamp2k(wt1 >> 10);
            break;
        }
        case 9: {
            i32 j = cr->b;
            cr->cxt = (pr->h[i] + pr->c8) * 32;
            i32 pq = 992;
            if (j < n) pq = p[j] + 992;
            if (pq < 0) pq = 0;
            if (pq > 1983) pq = 1983;
            i32 wt = pq & 63;
            pq >>= 6;
            i32 idx = (i32)cr->cxt + pq;
            i32 idx2 = idx + 1;
            if (idx >= 0 && idx2 < cr->cm_len) {
                i32 p1 = (i32)(cr->cm[idx] >> 10), p2 = (i32)(cr->cm[idx2] >> 10);
                p[i] = stretch((p1 * (64 - wt) + p2 * wt) >> 13);
            } else {
                p[i] = 0;
            }
            cr->cxt = (u32)idx + (u32)(wt >> 5);
            break;
        }
        default:
            p[i] = 0;
            break;
        }
    }
    return squash(p[n - 1]);
}

/* predictor.v:672-824 */
static void pred_update(pred_t *pr, i32 y) {
    i32 n = pr->n;
    i32 *p = pr->p;
    for (i32 i = 0; i < n; i++) {
        comp_t *cr = &pr->comp[i];
        switch (cr->ctype) {
        case 2: {
            i32 idx = (i32)cr->cxt & (cr->cm_len - 1);
            u32 pn = cr->cm[idx];
            i32 count = (i32)(pn & 0x3ff);
            i32 err = y * 32767 - (i32)(pn >> 17);
            i32 dt_val = count < 1024 ? g_dt[count] : g_dt[1023];
            i32 update = (err * dt_val) & -1024;
            i32 count_inc = count < cr->limit ? 1 : 0;
            cr->cm[idx] = (u32)((i32)pn + update + count_inc);
            break;
        }
        case 3: {
            i32 at = cr->c + (i32)(pr->hmap4 & 15);
            cr->ht[at] = (u8)st_next(cr->ht[at], y);
            u32 v = cr->cm[cr->cxt];
            cr->cm[cr->cxt] = (u32)((i32)v + ((y * 32767 - (i32)(v >> 8)) >> 2));
            break;
        }
        case 4: {
            if (cr->c != y) cr->a = 0;
            i32 idx = cr->limit & (cr->ht_len - 1);
            cr->ht[idx] = (u8)(((u32)cr->ht[idx] << 1) | (u32)y);
            cr->cxt++;
            if (cr->cxt >= 8) {
                cr->cxt = 0;
                cr->limit++;
                cr->limit &= (cr->ht_len - 1);
                if (cr->a == 0) {
                    u32 h = pr->h[i];
                    cr->b = cr->limit - (i32)cr->cm[(i32)h & (cr->cm_len - 1)];
                    if ((cr->b & (cr->ht_len - 1)) != 0) {
                        while (cr->a < 255) {
                            i32 idx1 = (cr->limit - cr->a - 1) & (cr->ht_len - 1);
                            i32 idx2 = (cr->limit - cr->a - cr->b - 1) & (cr->ht_len - 1);
                            if (cr->ht[idx1] != cr->ht[idx2]) break;
                            cr->a++;
                        }
                    }
                } else if (cr->a < 255) {
                    cr->a++;
                }
                cr->cm[(i32)pr->h[i] & (cr->cm_len - 1)] = (u32)cr->limit;
            }
            break;
        }
        case 6: {
            i32 j = (i32)cr->cm[0], k = (i32)cr->cm[1], rate = (i32)cr->cm[2];
            i32 err = ((y * 32767 - squash(p[i])) * rate) >> 5;
            if (j < n && k < n) {
                i32 w = cr->a16[cr->cxt];
                w += (err * (p[j] - p[k]) + (1 << 12)) >> 13;
                if (w < 0) w = 0;
                if (w > 65535) w = 65535;
                cr->a16[cr->cxt] = (u16)w;
            }
            break;
        }
        case 7: {
            i32 jj = cr->b, m = cr->limit, rate = cr->ht[0];
            i32 err = ((y * 32767 - squash(p[i])) * rate) >> 4;
            i32 idx = (i32)cr->cxt * m;
            for (i32 l = 0; l < m && (jj + l) < n; l++) {
                i32 wt = clamp512k((i32)cr->cm[idx + l] + ((err * p[jj + l] + (1 << 12)) >> 13));
                cr->cm[idx + l] = (u32)wt;
            }
            break;
        }
        case 8: {
            i32 j = cr->b;
            i32 err = y * 32767 - squash(p[i]);
            if (j < n) {
                i32 wt0 = clamp512k((i32)cr->cm[cr->cxt * 2] + ((err * p[j] + (1 << 12)) >> 13));
                i32 wt1 = clamp512k((i32)cr->cm[cr->cxt * 2 + 1] + ((err + 16) >> 5));
                cr->cm[cr->cxt * 2] = (u32)wt0;
                cr->cm[cr->cxt * 2 + 1] = (u32)wt1;
            }
            cr->ht[cr->c + (i32)(pr->hmap4 & 15)] = (u8)st_next((i32)cr->cxt, y);
            break;
        }
        case 9: {
            i32 idx = (i32)cr->cxt & (cr->cm_len - 1);
            u32 v = cr->cm[idx];
            i32 err = y * 32767 - (i32)(v >> 17);
            i32 count = (i32)v & 1023;
            if (count < cr->limit) v = (u32)((i32)v + ((err * (cr->limit - count) + (1 << 12)) >> 13) + 1);
            cr->cm[idx] = v;
            break;
        }
        default:
            break;
        }
    }
    pr->c8 = (pr->c8 << 1) | (u32)y;
    if (pr->c8 >= 256) {
        if (pr->z) {
            vm_run(pr->z, pr->c8 - 256);
            for (i32 i = 0; i < n && (u32)i < pr->z->hlen; i++) pr->h[i] = pr->z->h[i];
        }
        pr->hmap4 = 1;
        pr->c8 = 1;
    } else if (pr->c8 >= 16 && pr->c8 < 32) {
        pr->hmap4 = ((pr->hmap4 & 0xf) << 5) | ((u32)y << 4) | 1;
    } else {
        pr->hmap4 = (pr->hmap4 & 0x1f0) | (((pr->hmap4 & 0xf) * 2 + (u32)y) & 0xf);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Encoder (encoder.v) / Decoder (decoder.v)                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    u32 low, high;
    pred_t *pr;
    zo_buf *out;
} enc_t;

static void enc_init(enc_t *e, pred_t *pr, zo_buf *out) {
    e->low = 1, e->high = 0xFFFFFFFFu, e->pr = pr, e->out = out;
}
/* encoder.v:48-89 */
static void enc_encode(enc_t *e, i32 y, i32 p) {
    i32 pr = p < 0 ? 0 : (p > 65535 ? 65535 : p);
    u32 range_ = e->high - e->low;
    u32 mid = e->low + (u32)(((u64)range_ * (u64)pr) >> 16);
    if (y != 0)
        e->high = mid;
    else
        e->low = mid + 1;
    while ((e->high ^ e->low) < 0x1000000u) {
        if (e->out) buf_put(e->out, (int)(e->high >> 24));
        e->low <<= 8;
        e->high = (e->high << 8) | 0xFF;
        if (e->low == 0) e->low = 1;
    }
}
/* encoder.v:93-120 */
static void enc_compress(enc_t *e, i32 c) {
    if (!e->pr) return;
    if (c == -1) {
        enc_encode(e, 1, 0);
        return;
    }
    enc_encode(e, 0, 0);
    for (int i = 7; i >= 0; i--) {
        i32 y = (c >> i) & 1;
        i32 p = pred_predict(e->pr);
        enc_encode(e, y, p * 2 + 1);
        pred_update(e->pr, y);
    }
}
/* encoder.v:130-139 */
static void enc_flush(enc_t *e) {
    if (!e->out) return;
    buf_put(e->out, (int)(e->high >> 24));
    buf_put(e->out, (int)((e->high >> 16) & 255));
    buf_put(e->out, (int)((e->high >> 8) & 255));
    buf_put(e->out, (int)(e->high & 255));
}

typedef struct {
    const u8 *data;
    size_t len, pos;
} rdr_t;
static inline int rdr_get(rdr_t *r) { return r->pos < r->len ? r->data[r->pos++] : -1; }

typedef struct {
    u32 low, high, code;
    pred_t *pr;
    rdr_t *in;
} dec_t;

/* decoder.v:29-47 */
static void dec_init(dec_t *d, pred_t *pr, rdr_t *in) {
    d->pr = pr, d->in = in, d->low = 1, d->high = 0xFFFFFFFFu, d->code = 0;
    for (int i = 0; i < 4; i++) {
        int c = rdr_get(in);
        d->code = c < 0 ? (d->code << 8) : ((d->code << 8) | (u32)c);
    }
}
/* decoder.v:73-118 */
static i32 dec_decode(dec_t *d, i32 p) {
    i32 pr = p < 0 ? 0 : (p > 65535 ? 65535 : p);
    u32 range_ = d->high - d->low;
    u32 mid = d->low + (u32)(((u64)range_ * (u64)pr) >> 16);
    i32 y;
    if (d->code <= mid)
        y = 1, d->high = mid;
    else
        y = 0, d->low = mid + 1;
    while ((d->high ^ d->low) < 0x1000000u) {
        d->low <<= 8;
        d->high = (d->high << 8) | 0xFF;
        if (d->low == 0) d->low = 1;
        int c = rdr_get(d->in);
        d->code = c < 0 ? (d->code << 8) : ((d->code << 8) | (u32)c);
    }
    return y;
}
/* decoder.v:122-145 */
static i32 dec_decompress(dec_t *d) {
    if (!d->pr) return -1;
    if (dec_decode(d, 0) != 0) return -1;
    u32 c = 1;
    while (c < 256) {
        i32 p = pred_predict(d->pr);
        i32 y = dec_decode(d, p * 2 + 1);
        pred_update(d->pr, y);
        c = (c << 1) | (u32)y;
    }
    return (i32)c - 256;
}
/* decoder.v:151-196 */
static i32 dec_skip(dec_t *d) {
    if (!d->pr || d->pr->n == 0) return rdr_get(d->in);
    u32 curr = d->code;
    if (curr == 0) {
        int c = rdr_get(d->in);
        if (c < 0) return -1;
        curr = (u32)c;
    }
    int c = 0;
    while (curr != 0) {
        c = rdr_get(d->in);
        if (c < 0) return -1;
        curr = (curr << 8) | (u32)c;
    }
    for (;;) {
        c = rdr_get(d->in);
        if (c < 0) return -1;
        if (c != 0) break;
    }
    return c;
}

/* ------------------------------------------------------------------------------------------ */
/* Compressor (compressor.v)                                                                  */
/* ------------------------------------------------------------------------------------------ */
enum { CS_BLOCK = 0, CS_SEGMENT = 1, CS_START = 2 };
static const u8 k_locator[13] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83,
                                 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3};

struct zo_compressor {
    int state;
    zvm z;
    enc_t enc;
    pred_t pr;
    rdr_t in;
    int has_input;
    zo_buf out;
    sha1_t sha1;
    int level;
    u8 *store_buf;
    u32 store_size;
    int first_byte;
};

zo_compressor *zo_compressor_new(void) {
    zo_compressor *c = (zo_compressor *)calloc(1, sizeof(*c));
    c->state = CS_START;
    vm_new(&c->z);
    sha1_init(&c->sha1);
    c->level = 1;
    c->store_buf = (u8 *)malloc(65536 + 16);
    c->first_byte = 1;
    return c;
}
void zo_compressor_free(zo_compressor *c) {
    if (!c) return;
    pred_free(&c->pr);
    vm_free(&c->z);
    zo_buf_free(&c->out);
    free(c->store_buf);
    free(c);
}
void zo_compressor_set_input(zo_compressor *c, const u8 *data, size_t n) {
    c->in.data = data, c->in.len = n, c->in.pos = 0, c->has_input = 1;
}
zo_buf *zo_compressor_output(zo_compressor *c) { return &c->out; }

static void comp_start_block_common(zo_compressor *c, const u8 *hdr, int n) {
    /* compressor.v:90-148 */
    vm_clear(&c->z);
    vm_set_header(&c->z, hdr, n);
    vm_parse_geometry(&c->z);
    vm_inith(&c->z);
    vm_initp(&c->z);
    /* compressor.v:150-181 */
    zo_buf *o = &c->out;
    for (int i = 0; i < 13; i++) buf_put(o, k_locator[i]);
    buf_put(o, 0x7a), buf_put(o, 0x50), buf_put(o, 0x51);
    buf_put(o, (c->z.hlen_hdr >= 5 && c->z.header[4] != 0) ? 1 : 2);
    buf_put(o, 1);
    int hsize = (c->z.cend + 1) + (c->z.hend - c->z.hbegin + 1);
    buf_put(o, hsize & 0xFF), buf_put(o, (hsize >> 8) & 0xFF);
    for (int i = 0; i <= c->z.cend && i < c->z.hlen_hdr; i++) buf_put(o, c->z.header[i]);
    for (int i = c->z.hbegin; i <= c->z.hend && i < c->z.hlen_hdr; i++) buf_put(o, c->z.header[i]);
    /* compressor.v:184-187 */
    pred_init(&c->pr, &c->z);
    c->state = CS_BLOCK;
}
/* compressor.v:79-188 */
void zo_compressor_start_block(zo_compressor *c, int level) {
    if (c->state != CS_START) return;
    c->level = level;
    u8 hdr[80];
    int n = zo_level_header(level, hdr, (int)sizeof(hdr));
    comp_start_block_common(c, hdr, n);
}
void zo_compressor_start_block_header(zo_compressor *c, const u8 *hdr, int n) {
    if (c->state != CS_START) return;
    /* c.level drives the store-mode switch (compressor.v:265); a custom header with n>0
     * components behaves like any level >= 1. */
    c->level = (n >= 5 && hdr[4] != 0) ? 1 : 0;
    comp_start_block_common(c, hdr, n);
}
/* compressor.v:212-255 */
void zo_compressor_start_segment(zo_compressor *c, const char *filename, const char *comment) {
    if (c->state != CS_BLOCK) return;
    zo_buf *o = &c->out;
    buf_put(o, 1);
    for (const char *s = filename ? filename : ""; *s; s++) buf_put(o, (u8)*s);
    buf_put(o, 0);
    for (const char *s = comment ? comment : ""; *s; s++) buf_put(o, (u8)*s);
    buf_put(o, 0);
    buf_put(o, 0);
    enc_init(&c->enc, &c->pr, &c->out);
    sha1_init(&c->sha1);
    pred_reset(&c->pr);
    c->store_size = 0;
    c->first_byte = 1;
    c->state = CS_SEGMENT;
}
/* compressor.v:335-354 */
static void comp_flush_store(zo_compressor *c) {
    if (c->store_size == 0) return;
    zo_buf *o = &c->out;
    buf_put(o, (int)((c->store_size >> 24) & 0xFF)), buf_put(o, (int)((c->store_size >> 16) & 0xFF));
    buf_put(o, (int)((c->store_size >> 8) & 0xFF)), buf_put(o, (int)(c->store_size & 0xFF));
    for (u32 i = 0; i < c->store_size; i++) buf_put(o, c->store_buf[i]);
    c->store_size = 0;
}
/* compressor.v:259-332 */
int zo_compressor_compress(zo_compressor *c, int n) {
    if (c->state != CS_SEGMENT || !c->has_input) return 0;
    if (c->level == 0) {
        if (c->first_byte) {
            c->store_buf[c->store_size++] = 0;
            c->first_byte = 0;
        }
        int count = 0;
        while (count < n) {
            int ch = rdr_get(&c->in);
            if (ch < 0) return 0;
            sha1_put(&c->sha1, ch);
            c->store_buf[c->store_size++] = (u8)ch;
            if (c->store_size >= 65536) comp_flush_store(c);
            count++;
        }
        return 1;
    }
    if (c->first_byte) {
        enc_compress(&c->enc, 0);
        c->first_byte = 0;
    }
    int count = 0;
    while (count < n) {
        int ch = rdr_get(&c->in);
        if (ch < 0) return 0;
        sha1_put(&c->sha1, ch);
        enc_compress(&c->enc, ch);
        count++;
    }
    return 1;
}
/* compressor.v:357-399 */
void zo_compressor_end_segment(zo_compressor *c) {
    if (c->state != CS_SEGMENT) return;
    zo_buf *o = &c->out;
    if (c->level == 0 || c->pr.n == 0) {
        comp_flush_store(c);
    } else {
        enc_compress(&c->enc, -1);
        enc_flush(&c->enc);
    }
    buf_put(o, 0), buf_put(o, 0), buf_put(o, 0), buf_put(o, 0);
    u8 hash[20];
    sha1_result(&c->sha1, hash);
    buf_put(o, 253);
    for (int i = 0; i < 20; i++) buf_put(o, hash[i]);
    c->state = CS_BLOCK;
}
/* compressor.v:402-413 */
void zo_compressor_end_block(zo_compressor *c) {
    if (c->state != CS_BLOCK) return;
    buf_put(&c->out, 0xFF);
    c->state = CS_START;
}

/* ------------------------------------------------------------------------------------------ */
/* Decompresser (decompressor.v)                                                              */
/* ------------------------------------------------------------------------------------------ */
enum { DS_BLOCK = 0, DS_SEGMENT = 1, DS_FILENAME = 2, DS_START = 3 };

struct zo_decompresser {
    int state;
    zvm z;
    dec_t dec;
    pred_t pr;
    rdr_t in;
    int has_input;
    zo_buf out;
    sha1_t sha1;
    zo_buf filename, comment;
    u32 store_count;
    int first_seg;
    int pp_state; /* PostProcessor.state (decompressor.v:56-152); PROG states are out of scope */
    int last_sha1_ok;
};

zo_decompresser *zo_decompresser_new(void) {
    zo_decompresser *d = (zo_decompresser *)calloc(1, sizeof(*d));
    d->state = DS_START;
    vm_new(&d->z);
    sha1_init(&d->sha1);
    d->first_seg = 1;
    d->last_sha1_ok = -1;
    buf_put(&d->filename, 0), d->filename.len = 0;
    buf_put(&d->comment, 0), d->comment.len = 0;
    return d;
}
void zo_decompresser_free(zo_decompresser *d) {
    if (!d) return;
    pred_free(&d->pr);
    vm_free(&d->z);
    zo_buf_free(&d->out), zo_buf_free(&d->filename), zo_buf_free(&d->comment);
    free(d);
}
void zo_decompresser_set_input(zo_decompresser *d, const u8 *data, size_t n) {
    d->in.data = data, d->in.len = n, d->in.pos = 0, d->has_input = 1;
}
size_t zo_decompresser_input_pos(const zo_decompresser *d) { return d->in.pos; }
zo_buf *zo_decompresser_output(zo_decompresser *d) { return &d->out; }
const char *zo_decompresser_filename(const zo_decompresser *d) { return (const char *)d->filename.data; }
const char *zo_decompresser_comment(const zo_decompresser *d) { return (const char *)d->comment.data; }
int zo_decompresser_last_sha1_ok(const zo_decompresser *d) { return d->last_sha1_ok; }

/* decompressor.v:219-346 */
int zo_decompresser_find_block(zo_decompresser *d) {
    if (!d->has_input) return 0;
    u32 h1 = 0x3D49B113u, h2 = 0x29EB7F93u, h3 = 0x2614BE13u, h4 = 0x3828EB13u;
    for (;;) {
        int c = rdr_get(&d->in);
        if (c < 0) return 0;
        h1 = h1 * 12 + (u32)c, h2 = h2 * 20 + (u32)c, h3 = h3 * 28 + (u32)c, h4 = h4 * 44 + (u32)c;
        if (h1 == 0xB16B88F1u && h2 == 0xFF5376F1u && h3 == 0x72AC5BF1u && h4 == 0x2F909AF1u) break;
    }
    int level = rdr_get(&d->in);
    if (level < 0 || (level != 1 && level != 2)) return 0;
    int block_type = rdr_get(&d->in);
    if (block_type != 1) return 0;
    int lo = rdr_get(&d->in), hi = rdr_get(&d->in);
    if (lo < 0 || hi < 0) return 0;
    int hsize = lo + hi * 256;

    vm_free(&d->z);
    vm_new(&d->z);
    u8 *hdr = (u8 *)malloc((size_t)hsize + 2048);
    int hl = 0;
    int ok = 0;
    do {
        int bad = 0;
        for (int i = 0; i < 5; i++) {
            int b = rdr_get(&d->in);
            if (b < 0) { bad = 1; break; }
            hdr[hl++] = (u8)b;
        }
        if (bad) break;
        int n = hdr[4];
        for (int i = 0; i < n && !bad; i++) {
            int ctype = rdr_get(&d->in);
            if (ctype < 0 || ctype >= 10) { bad = 1; break; }
            hdr[hl++] = (u8)ctype;
            for (int j = 1; j < k_compsize[ctype]; j++) {
                int b = rdr_get(&d->in);
                if (b < 0) { bad = 1; break; }
                hdr[hl++] = (u8)b;
            }
        }
        if (bad) break;
        if (rdr_get(&d->in) != 0) break;
        hdr[hl++] = 0;
        d->z.cend = hl - 1;
        d->z.hbegin = hl;
        int hcomp_len = hsize - hl;
        for (int i = 0; i < hcomp_len; i++) {
            int b = rdr_get(&d->in);
            if (b < 0) { bad = 1; break; }
            hdr[hl++] = (u8)b;
        }
        if (bad) break;
        ok = 1;
    } while (0);
    if (!ok) {
        free(hdr);
        return 0;
    }
    vm_set_header(&d->z, hdr, hl);
    free(hdr);
    d->z.hend = hl - 1;
    vm_inith(&d->z);
    vm_initp(&d->z);
    pred_init(&d->pr, &d->z);
    d->state = DS_BLOCK;
    return 1;
}

static void str_reset(zo_buf *b) { b->len = 0; }
static void str_term(zo_buf *b) { buf_put(b, 0), b->len--; }

/* decompressor.v:350-429 */
int zo_decompresser_find_filename(zo_decompresser *d) {
    if (d->state != DS_BLOCK || !d->has_input) return 0;
    int marker = rdr_get(&d->in);
    if (marker < 0) return 0;
    if (marker == 0xFF) {
        d->state = DS_START;
        return 0;
    }
    str_reset(&d->filename);
    for (;;) {
        int c = rdr_get(&d->in);
        if (c < 0) return 0;
        if (c == 0) break;
        if (c == 0xFF) {
            d->state = DS_START;
            return 0;
        }
        buf_put(&d->filename, c);
    }
    str_term(&d->filename);
    str_reset(&d->comment);
    for (;;) {
        int c = rdr_get(&d->in);
        if (c < 0) return 0;
        if (c == 0) break;
        buf_put(&d->comment, c);
    }
    str_term(&d->comment);
    if (rdr_get(&d->in) < 0) return 0;
    if (d->pr.n > 0) {
        pred_reset(&d->pr);
        dec_init(&d->dec, &d->pr, &d->in);
    }
    sha1_init(&d->sha1);
    d->store_count = 0;
    d->first_seg = 1;
    d->last_sha1_ok = -1;
    d->state = DS_SEGMENT;
    return 1;
}

static void dec_emit(zo_decompresser *d, int b) {
    sha1_put(&d->sha1, b);
    buf_put(&d->out, b);
}

/* decompressor.v:518-587 */
static int decompress_store(zo_decompresser *d, int n) {
    int count = 0;
    int limit = n < 0 ? 0x7FFFFFFF : n;
    while (count < limit) {
        if (d->store_count == 0) {
            int b0 = rdr_get(&d->in), b1 = rdr_get(&d->in), b2 = rdr_get(&d->in), b3 = rdr_get(&d->in);
            if (b0 < 0 || b1 < 0 || b2 < 0 || b3 < 0) return 0;
            d->store_count = ((u32)b0 << 24) | ((u32)b1 << 16) | ((u32)b2 << 8) | (u32)b3;
            if (d->store_count == 0) return 0;
            if (d->first_seg) {
                if (rdr_get(&d->in) < 0) return 0;
                d->store_count--;
                d->first_seg = 0;
                if (d->store_count == 0) continue;
            }
        }
        int c = rdr_get(&d->in);
        if (c < 0) return 0;
        dec_emit(d, c);
        d->store_count--;
        count++;
    }
    return 1;
}

/* PostProcessor.write, states 0 and 1 only (decompressor.v:56-82).  PROG (c==1) is never
 * emitted by the reference compressor; the oracle treats it as PASS-through of nothing. */
static int pp_write(zo_decompresser *d, int c, int *emitted) {
    *emitted = -1;
    if (d->pp_state == 0) {
        if (c < 0) return d->pp_state;
        d->pp_state = c + 1;
        if (d->pp_state > 2) d->pp_state = 1;
    } else if (d->pp_state == 1) {
        if (c >= 0) *emitted = c;
    }
    return d->pp_state;
}

/* decompressor.v:443-515 */
int zo_decompresser_decompress(zo_decompresser *d, int n) {
    if (d->state != DS_SEGMENT) return 0;
    if (d->pr.n == 0) return decompress_store(d, n);
    if (d->first_seg) {
        d->pp_state = 0;
        d->first_seg = 0;
    }
    int em;
    while ((d->pp_state & 3) != 1) {
        i32 c = dec_decompress(&d->dec);
        if (c < 0) return 0;
        pp_write(d, c, &em);
        if (d->pp_state == 2) return 0; /* PROG: out of scope */
    }
    int count = 0;
    int limit = n < 0 ? 0x7FFFFFFF : n;
    while (count < limit) {
        i32 c = dec_decompress(&d->dec);
        pp_write(d, c, &em);
        if (c < 0) return 0;
        if (em >= 0) {
            dec_emit(d, em);
            count++;
        }
    }
    return 1;
}

/* decompressor.v:590-635 */
void zo_decompresser_read_segment_end(zo_decompresser *d) {
    if (d->state != DS_SEGMENT) return;
    int marker = (d->pr.n > 0) ? dec_skip(&d->dec) : rdr_get(&d->in);
    if (marker == 253) {
        u8 stored[20] = {0}, computed[20];
        for (int i = 0; i < 20; i++) {
            int c = rdr_get(&d->in);
            if (c >= 0) stored[i] = (u8)c;
        }
        sha1_result(&d->sha1, computed);
        d->last_sha1_ok = memcmp(stored, computed, 20) == 0;
    }
    d->state = DS_BLOCK;
}

/* ------------------------------------------------------------------------------------------ */
/* raw coder helpers (zpaq_test.v:430-527 shape)                                              */
/* ------------------------------------------------------------------------------------------ */
size_t zo_raw_encode(const u8 *hdr, int hdr_len, const u8 *data, size_t n, int with_pp, u8 **out) {
    zvm z;
    pred_t pr;
    enc_t e;
    zo_buf o = {0};
    vm_new(&z);
    memset(&pr, 0, sizeof(pr));
    vm_set_header(&z, hdr, hdr_len);
    vm_parse_geometry(&z);
    vm_inith(&z), vm_initp(&z);
    pred_init(&pr, &z);
    enc_init(&e, &pr, &o);
    if (with_pp) enc_compress(&e, 0);
    for (size_t i = 0; i < n; i++) enc_compress(&e, data[i]);
    enc_compress(&e, -1);
    enc_flush(&e);
    pred_free(&pr), vm_free(&z);
    *out = o.data;
    return o.len;
}
size_t zo_raw_decode(const u8 *hdr, int hdr_len, const u8 *code, size_t n, u8 **out) {
    zvm z;
    pred_t pr;
    dec_t d;
    rdr_t r = {code, n, 0};
    zo_buf o = {0};
    vm_new(&z);
    memset(&pr, 0, sizeof(pr));
    vm_set_header(&z, hdr, hdr_len);
    vm_parse_geometry(&z);
    vm_inith(&z), vm_initp(&z);
    pred_init(&pr, &z);
    dec_init(&d, &pr, &r);
    for (;;) {
        i32 c = dec_decompress(&d);
        if (c < 0) break;
        buf_put(&o, c);
    }
    pred_free(&pr), vm_free(&z);
    *out = o.data;
    return o.len;
}

/* ------------------------------------------------------------------------------------------ */
/* convenience drivers                                                                        */
/* ------------------------------------------------------------------------------------------ */
size_t zo_compress_block(int level, const u8 *hdr, int hdr_len, const u8 *data, size_t n,
                         const char *filename, const char *comment, u8 **out) {
    zo_compressor *c = zo_compressor_new();
    zo_compressor_set_input(c, data, n);
    if (hdr)
        zo_compressor_start_block_header(c, hdr, hdr_len);
    else
        zo_compressor_start_block(c, level);
    zo_compressor_start_segment(c, filename, comment);
    /* cmd/main.v:305: for comp.compress(65536) {} */
    while (zo_compressor_compress(c, 65536)) {
    }
    zo_compressor_end_segment(c);
    zo_compressor_end_block(c);
    size_t len = c->out.len;
    *out = c->out.data;
    c->out.data = NULL, c->out.len = c->out.cap = 0;
    zo_compressor_free(c);
    return len;
}

size_t zo_decompress_archive(const u8 *arc, size_t n, u8 **out, int *n_segments, int *n_bad_sha) {
    zo_decompresser *d = zo_decompresser_new();
    zo_decompresser_set_input(d, arc, n);
    int segs = 0, bad = 0;
    /* cmd/main.v:349-401 */
    while (zo_decompresser_find_block(d)) {
        while (zo_decompresser_find_filename(d)) {
            while (zo_decompresser_decompress(d, 65536)) {
            }
            zo_decompresser_read_segment_end(d);
            segs++;
            if (d->last_sha1_ok == 0) bad++;
        }
    }
    size_t len = d->out.len;
    *out = d->out.data;
    d->out.data = NULL, d->out.len = d->out.cap = 0;
    zo_decompresser_free(d);
    if (n_segments) *n_segments = segs;
    if (n_bad_sha) *n_bad_sha = bad;
    return len;
}

typedef struct {
    int level, decompress;
    const u8 *in;
    const u64 *in_off;
    int n_blocks;
    u8 **res;
    u64 *res_len;
    volatile int *next;
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *j = (mt_job *)arg;
    for (;;) {
        int k = __sync_fetch_and_add(j->next, 1);
        if (k >= j->n_blocks) break;
        const u8 *src = j->in + j->in_off[k];
        size_t n = (size_t)(j->in_off[k + 1] - j->in_off[k]);
        if (j->decompress) {
            j->res_len[k] = zo_decompress_archive(src, n, &j->res[k], NULL, NULL);
        } else {
            char comment[32];
            size_t v = n;
            int pos = 0;
            char tmp[24];
            do tmp[pos++] = (char)('0' + v % 10), v /= 10; while (v);
            int q = 0;
            while (pos) comment[q++] = tmp[--pos];
            memcpy(comment + q, " bytes", 7);
            j->res_len[k] = zo_compress_block(j->level, NULL, 0, src, n, "", comment, &j->res[k]);
        }
    }
    return NULL;
}

static int run_mt(int level, int decompress, const u8 *in, const u64 *in_off, int n_blocks, u8 *out,
                  u64 out_cap, u64 *out_off, u64 *out_need, int threads) {
    ensure_tables();
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    u8 **res = (u8 **)calloc((size_t)n_blocks + 1, sizeof(u8 *));
    u64 *res_len = (u64 *)calloc((size_t)n_blocks + 1, sizeof(u64));
    volatile int next = 0;
    mt_job job = {level, decompress, in, in_off, n_blocks, res, res_len, &next};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, mt_worker, &job);
    mt_worker(&job);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    u64 total = 0;
    for (int k = 0; k < n_blocks; k++) total += res_len[k];
    if (out_need) *out_need = total;
    int rc = 0;
    if (total > out_cap) {
        rc = -1;
    } else {
        u64 pos = 0;
        for (int k = 0; k < n_blocks; k++) {
            out_off[k] = pos;
            memcpy(out + pos, res[k], (size_t)res_len[k]);
            pos += res_len[k];
        }
        out_off[n_blocks] = pos;
    }
    for (int k = 0; k < n_blocks; k++) free(res[k]);
    free(res), free(res_len);
    return rc;
}

int zo_compress_blocks_mt(int level, const u8 *in, const u64 *in_off, int n_blocks, u8 *out,
                          u64 out_cap, u64 *out_off, u64 *out_need, int threads) {
    return run_mt(level, 0, in, in_off, n_blocks, out, out_cap, out_off, out_need, threads);
}
int zo_decompress_blocks_mt(const u8 *arc, const u64 *arc_off, int n_blocks, u8 *out, u64 out_cap,
                            u64 *out_off, u64 *out_need, int threads) {
    return run_mt(0, 1, arc, arc_off, n_blocks, out, out_cap, out_off, out_need, threads);
}

// model.cpp -- header parsing, workspace layout, constant tables (host side of libzpaqgpu).
#include "model.h"

#include <algorithm>
#include <cstring>
#include <mutex>

#include "../../include/zpaqgpu.h"

namespace zg {

// ---------------------------------------------------------------------------------------------
// Lookup tables.  squash/stretch are produced by the reference at start-up from hand-rolled f64
// series (predictor.v:21-96, :169-190) that do not converge for large arguments; the operation
// order below is the reference's and this file is compiled with -ffp-contract=off.  The device
// only ever sees the finished integer tables.
// ---------------------------------------------------------------------------------------------
namespace {

double series_exp(double x) {                       // predictor.v:52-70
    if (x < -20.0) return 0.0;
    if (x > 20.0) return 485165195.4;
    double acc = 1.0, t = 1.0;
    for (int k = 1; k < 40; ++k) {
        t *= x / double(k);
        acc += t;
        if (t < 1e-15 && t > -1e-15) break;
    }
    return acc;
}

double series_ln(double x) {                        // predictor.v:169-190
    if (x <= 0.0) return -20.0;
    if (x > 1e9) return 20.0;
    const double y = (x - 1.0) / (x + 1.0), y2 = y * y;
    double acc = y, t = y;
    for (int k = 1; k < 50; ++k) {
        t *= y2;
        acc += t / double(2 * k + 1);
        if (t < 1e-15 && t > -1e-15) break;
    }
    return 2.0 * acc;
}

// The bit-history automaton of statetable.v:15-57 is libzpaq's `sns` table; it is regenerated
// from libzpaq's published construction (bounded (n0,n1) count pairs, discounting of the
// opposite count) instead of being carried as a literal.
struct StateGen {
    static int states(int n0, int n1) {
        static const int cap[6] = {20, 48, 15, 8, 6, 5};
        if (n0 < n1) return states(n1, n0);
        if (n0 < 0 || n1 < 0 || n1 >= 6 || n0 > cap[n1]) return 0;
        return 1 + int(n1 > 0 && n0 + n1 <= 17);
    }
    static int discounted(int c) {
        return int(c >= 1) + int(c >= 2) + int(c >= 3) + int(c >= 4) + int(c >= 5) + int(c >= 7) +
               int(c >= 8);
    }
    static void step(int &n0, int &n1, int bit) {
        if (n0 < n1) return step(n1, n0, 1 - bit);
        if (bit) {
            ++n1;
            n0 = discounted(n0);
        } else {
            ++n0;
            n1 = discounted(n1);
        }
        while (!states(n0, n1)) {
            if (n1 < 2) {
                --n0;
            } else {
                n0 = (n0 * (n1 - 1) + (n1 / 2)) / n1;
                --n1;
            }
        }
    }
    static void build(uint8_t *ns) {
        constexpr int N = 50;
        static uint8_t id[N][N][2];
        std::memset(id, 0, sizeof(id));
        int next_id = 0;
        for (int total = 0; total < N; ++total)
            for (int n1 = 0; n1 <= total; ++n1) {
                const int n0 = total - n1, k = states(n0, n1);
                if (!k) continue;
                id[n0][n1][0] = uint8_t(next_id);
                id[n0][n1][1] = uint8_t(next_id + k - 1);
                next_id += k;
            }
        std::memset(ns, 0, 1024);
        for (int n0 = 0; n0 < N; ++n0)
            for (int n1 = 0; n1 < N; ++n1)
                for (int v = 0; v < states(n0, n1); ++v) {
                    const int s = id[n0][n1][v];
                    int a = n0, b = n1;
                    step(a, b, 0);
                    ns[s * 4 + 0] = id[a][b][0];
                    a = n0, b = n1;
                    step(a, b, 1);
                    ns[s * 4 + 1] = id[a][b][1];
                    ns[s * 4 + 2] = uint8_t(n0);
                    ns[s * 4 + 3] = uint8_t(n1);
                }
    }
};

Tables g_tables;
std::once_flag g_tables_once;

void build_tables() {
    Tables &t = g_tables;
    std::memset(&t, 0, sizeof(t));
    for (int d = -2047; d <= 2047; ++d) {           // predictor.v:21-49
        double x = double(d) / 64.0;
        if (x < -20.0) x = -20.0;
        if (x > 20.0) x = 20.0;
        double e;
        if (x >= 0) {
            e = 1.0 / (1.0 + series_exp(-x));
        } else {
            const double u = series_exp(x);
            e = u / (1.0 + u);
        }
        const int v = int(32767.0 * e + 0.5);
        t.squash[d + 2047] = v < 1 ? 1 : (v > 32767 ? 32767 : v);
    }
    for (int i = 0; i < 32768; ++i) {               // predictor.v:73-96
        const double p = double(i) / 32767.0;
        if (p <= 0.0) {
            t.stretch[i] = -2047;
        } else if (p >= 1.0) {
            t.stretch[i] = 2047;
        } else {
            const int v = int(series_ln(p / (1.0 - p)) * 64.0);
            t.stretch[i] = v < -2047 ? -2047 : (v > 2047 ? 2047 : v);
        }
    }
    for (int i = 0; i < 256; ++i) t.dt2k[i] = 2048 - 2048 / (i + 1);          // predictor.v:99-106
    for (int i = 0; i < 1024; ++i) t.dt[i] = (1 << 17) / (i * 2 + 3) * 2;     // predictor.v:109-166
    StateGen::build(t.ns);
}

constexpr int kCompSize[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};                 // types.v:74-85

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

}  // namespace

const Tables &tables() {
    std::call_once(g_tables_once, build_tables);
    return g_tables;
}

int h_squash(int d) {                               // predictor.v:193-202
    int idx = d + 2047;
    if (idx < 0) idx = 0;
    if (idx >= 4094) idx = 4093;
    return tables().squash[idx];
}
int h_stretch(int p) {                              // predictor.v:205-214
    int idx = p < 1 ? 1 : (p >= 32768 ? 32767 : p);
    return tables().stretch[idx];
}
int h_clamp512k(int x) { return x < -262144 ? -262144 : (x > 262143 ? 262143 : x); }
int st_cminit(int s) {                              // statetable.v:90-100
    if (s < 0 || s >= 256) return 1 << 22;
    const uint32_t n0 = tables().ns[s * 4 + 2], n1 = tables().ns[s * 4 + 3];
    return int(((n1 * 2 + 1) << 22) / (n0 + n1 + 1));
}

// ---------------------------------------------------------------------------------------------
// The five predefined models (levels.v:40-375), generated from their parameters.
// ---------------------------------------------------------------------------------------------
std::vector<uint8_t> level_header(int level) {
    std::vector<uint8_t> h;
    if (level == 0) return std::vector<uint8_t>(7, 0);
    if (level < 2 || level > 5) {                   // level 1, and the fallback of levels.v:34
        // ICM16 + ISSE19; HCOMP: *b=a a=0 d=0 hash b-- hash *d=a d++ b-- hash b-- hash *d=a halt
        h = {1, 2, 0, 0, 2, C_ICM, 16, C_ISSE, 19, 0, 0};
        const uint8_t prog[] = {96, 4, 28, 59, 10, 59, 112, 25, 10, 59, 10, 59, 112, 56, 0};
        h.insert(h.end(), prog, prog + sizeof(prog));
        return h;
    }
    const int idx = level - 2;
    const uint8_t hh[4] = {9, 10, 12, 14}, bits[4] = {16, 18, 20, 22}, chain[4] = {2, 4, 5, 7};
    const bool mix = level >= 4;
    const int n = 1 + chain[idx] + int(mix);
    h = {hh[idx], bits[idx], 0, 0, uint8_t(n), C_ICM, bits[idx]};
    for (int i = 0; i < chain[idx]; ++i) {
        h.push_back(C_ISSE), h.push_back(bits[idx]), h.push_back(uint8_t(i));
    }
    if (mix) {
        const uint8_t mx[] = {C_MIX2, uint8_t(level == 4 ? 16 : 18), uint8_t(chain[idx] - 1),
                              chain[idx], 24, 255};
        h.insert(h.end(), mx, mx + 6);
    }
    h.push_back(0);
    // b=c c-- *c=a d=0, then one "hash *d=a" per component separated by d++, halt
    const uint8_t head[] = {74, 18, 104, 95, 0};
    h.insert(h.end(), head, head + 5);
    for (int i = 0; i < n; ++i) {
        h.push_back(59), h.push_back(112);
        if (i + 1 < n) h.push_back(25);
    }
    h.push_back(56), h.push_back(0), h.push_back(0);
    return h;
}

// ---------------------------------------------------------------------------------------------
// Component table + workspace layout (predictor.v:292-470)
// ---------------------------------------------------------------------------------------------
namespace {

uint32_t add_image(Model &m, const std::vector<uint32_t> &img) {
    const uint32_t at = uint32_t(m.image.size());
    m.image.insert(m.image.end(), img.begin(), img.end());
    return at;
}

int build_layout(Model &m, bool paged) {
    const std::vector<uint8_t> &hd = m.header;
    const int len = int(hd.size());
    std::vector<CompDesc> &comps_out = paged ? m.comps_paged : m.comps;
    std::vector<FillRegion> &fills_out = paged ? m.fills_paged : m.fills;
    comps_out.clear();
    fills_out.clear();
    if (!paged) m.image.clear();
    m.n = len >= 5 ? hd[4] : 0;
    // a hash table of `bytes` occupies its page table in the paged layout
    auto ht_resident = [&](uint64_t bytes) {
        return paged ? std::max<uint64_t>(4, bytes / kPageBytes * 4) : bytes;
    };
    uint64_t ws = 0;
    auto reserve = [&](uint64_t bytes) {
        ws = align_up(ws, 256);
        const uint64_t at = ws;
        ws += bytes;
        return at;
    };
    comps_out.assign(size_t(m.n), CompDesc{});
    int cp = 5;
    for (int i = 0; i < m.n && cp < m.cend; ++i) {
        CompDesc &c = comps_out[size_t(i)];
        const int type = hd[size_t(cp)];
        c.type = type;
        const int need = (type >= 0 && type < 10) ? kCompSize[type] : 1;
        if (cp + (need ? need : 1) > len) {
            m.error = "component parameters run past the header";
            return ZPAQGPU_E_FORMAT;
        }
        auto P = [&](int k) { return int(hd[size_t(cp + k)]); };
        switch (type) {
        case C_CONS:
            c.a = P(1);
            cp += 2;
            break;
        case C_CM: {
            c.a = P(1);
            c.limit = P(2) * 4;
            if (c.a > 26) return m.error = "CM sizebits > 26", ZPAQGPU_E_UNSUPPORTED;
            c.cm_len = 1u << c.a;
            c.cm_off = reserve(uint64_t(c.cm_len) * 4);
            fills_out.push_back({c.cm_off, c.cm_len, add_image(m, {0x80000000u}), 1});
            cp += 3;
            break;
        }
        case C_ICM: {
            c.a = P(1);
            if (c.a > 24) return m.error = "ICM sizebits > 24", ZPAQGPU_E_UNSUPPORTED;
            c.cm_len = 256;
            c.cm_off = reserve(256 * 4);
            c.ht_len = 16u << (c.a + 2);
            c.ht_off = reserve(ht_resident(c.ht_len));
            std::vector<uint32_t> img(256);
            for (int j = 0; j < 256; ++j) img[size_t(j)] = uint32_t(st_cminit(j));
            fills_out.push_back({c.cm_off, 256, add_image(m, img), 256});
            cp += 2;
            break;
        }
        case C_MATCH:
            c.a = P(1), c.b = P(2);
            if (c.a > 26 || c.b > 28) return m.error = "MATCH bits too large", ZPAQGPU_E_UNSUPPORTED;
            c.cm_len = 1u << c.a;
            c.cm_off = reserve(uint64_t(c.cm_len) * 4);
            c.ht_len = 1u << c.b;
            c.ht_off = reserve(c.ht_len);
            cp += 3;
            break;
        case C_AVG:
            c.a = P(1), c.b = P(2), c.c = P(3);
            cp += 4;
            break;
        case C_MIX2: {
            c.a = P(1);
            if (c.a > 26) return m.error = "MIX2 sizebits > 26", ZPAQGPU_E_UNSUPPORTED;
            c.b = P(2);
            c.c = 1 << c.a;
            c.p[0] = uint32_t(P(2)), c.p[1] = uint32_t(P(3)), c.p[2] = uint32_t(P(4)), c.p[3] = uint32_t(P(5));
            c.a16_len = 1u << c.a;
            const uint64_t words = (uint64_t(c.a16_len) * 2 + 3) / 4;
            c.a16_off = reserve(words * 4);
            fills_out.push_back({c.a16_off, words, add_image(m, {0x80008000u}), 1});
            cp += 6;
            break;
        }
        case C_MIX: {
            c.a = P(1);
            if (c.a > 22) return m.error = "MIX sizebits > 22", ZPAQGPU_E_UNSUPPORTED;
            const int size = 1 << c.a, j = P(2), mm = P(3);
            if (mm == 0) return m.error = "MIX with m=0 (the reference divides by zero)", ZPAQGPU_E_FORMAT;
            c.b = j, c.c = size, c.limit = mm;
            c.p[0] = uint32_t(P(4)), c.p[1] = uint32_t(P(5));
            c.cm_len = uint32_t(size) * uint32_t(mm);
            c.cm_off = reserve(uint64_t(c.cm_len) * 4);
            fills_out.push_back({c.cm_off, c.cm_len, add_image(m, {uint32_t(65536 / mm) << 8}), 1});
            cp += 6;
            break;
        }
        case C_ISSE: {
            c.a = P(1), c.b = P(2);
            if (c.a > 24) return m.error = "ISSE sizebits > 24", ZPAQGPU_E_UNSUPPORTED;
            c.cm_len = 512;
            c.cm_off = reserve(512 * 4);
            c.ht_len = 16u << (c.a + 2);
            c.ht_off = reserve(ht_resident(c.ht_len));
            std::vector<uint32_t> img(512);
            for (int k = 0; k < 256; ++k) {
                img[size_t(k) * 2] = 1u << 15;
                img[size_t(k) * 2 + 1] = uint32_t(h_clamp512k(h_stretch(st_cminit(k) >> 8) * 1024));
            }
            fills_out.push_back({c.cm_off, 512, add_image(m, img), 512});
            cp += 3;
            break;
        }
        case C_SSE: {
            c.a = P(1), c.b = P(2);
            if (c.a > 21) return m.error = "SSE sizebits > 21", ZPAQGPU_E_UNSUPPORTED;
            const int size = 1 << c.a, start = P(3);
            c.limit = P(4) * 4;
            c.cm_len = uint32_t(size) * 32;
            c.cm_off = reserve(uint64_t(c.cm_len) * 4);
            std::vector<uint32_t> img(32);
            for (int k = 0; k < 32; ++k) img[size_t(k)] = (uint32_t(h_squash(k * 64 - 992)) << 17) | uint32_t(start);
            fills_out.push_back({c.cm_off, c.cm_len, add_image(m, img), 32});
            cp += 5;
            break;
        }
        default:
            cp += 1;
            break;
        }
    }
    // evaluation levels (predictor.v:536-668 evaluates in index order: only inputs j < i are of this bit)
    for (int i = 0; i < m.n; ++i) {
        CompDesc &c = comps_out[size_t(i)];
        int deepest = -1;
        auto dep = [&](int j) {
            if (j >= 0 && j < i) deepest = std::max(deepest, int(comps_out[size_t(j)].level));
        };
        switch (c.type) {
        case C_AVG: dep(c.a), dep(c.b); break;
        case C_MIX2: dep(int(c.p[0])), dep(int(c.p[1])); break;
        case C_ISSE: case C_SSE: dep(c.b); break;
        case C_MIX:
            for (int l = 0; l < c.limit; ++l) dep(c.b + l);
            break;
        default: break;
        }
        const bool reads_others = c.type == C_AVG || c.type == C_MIX2 || c.type == C_ISSE || c.type == C_SSE ||
                                  c.type == C_MIX;
        c.level = reads_others ? deepest + 1 + (deepest < 0 ? 1 : 0) : 0;
    }
    // ZPAQL memory (zpaql.v:74-96): allocated only for 0 < bits < 32
    const int hh = len >= 2 ? hd[0] : 0, hm = len >= 2 ? hd[1] : 0;
    if (hh > 26 || hm > 30) return m.error = "H/M array too large", ZPAQGPU_E_UNSUPPORTED;
    if (paged) {  // chain kernels evaluate HCOMP in closed form: no ZPAQL memory in this layout
        m.ws_bytes_paged = align_up(ws, 256);
        return ZPAQGPU_OK;
    }
    m.h_len = hh > 0 ? 1u << hh : 0;
    m.m_len = hm > 0 ? 1u << hm : 0;
    m.h_off = reserve(uint64_t(m.h_len) * 4);
    m.m_off = reserve(m.m_len);
    m.r_off = reserve(256 * 4);
    m.rt_off = reserve(uint64_t(m.n > 0 ? m.n : 1) * 32);
    m.ws_bytes = align_up(ws, 256);
    return ZPAQGPU_OK;
}

void detect_shape(Model &m) {
    m.is_chain = false, m.n_isse = 0, m.has_mix2 = false, m.ctx_mode = CTX_VM, m.n_hash = 0;
    const int n = m.n;
    if (n < 1 || n > 9 || m.comps[0].type != C_ICM) return;
    int i = 1;
    while (i < n && m.comps[size_t(i)].type == C_ISSE && m.comps[size_t(i)].b == i - 1) ++i;
    const int n_isse = i - 1;
    bool mix2 = false;
    if (i == n - 1 && m.comps[size_t(i)].type == C_MIX2 && n_isse >= 2 &&
        int(m.comps[size_t(i)].p[0]) == n - 3 && int(m.comps[size_t(i)].p[1]) == n - 2) {
        mix2 = true;
        ++i;
    }
    if (i != n || n_isse > 7) return;
    // context program
    const uint8_t *prog = m.header.data() + m.hbegin;
    const int plen = m.hend - m.hbegin;
    static const uint8_t m1[] = {96, 4, 28, 59, 10, 59, 112, 25, 10, 59, 10, 59, 112, 56};
    int mode = CTX_VM, n_hash = 0;
    if (plen == int(sizeof(m1)) && std::memcmp(prog, m1, sizeof(m1)) == 0 && m.m_len == 4 && m.h_len >= 2) {
        mode = CTX_M1;
        n_hash = 2;
    } else if (plen >= 8 && prog[0] == 74 && prog[1] == 18 && prog[2] == 104 && prog[3] == 95 &&
               prog[4] == 0 && m.m_len >= 2) {
        int k = 5, rounds = 0;
        bool ok = true;
        for (;;) {
            if (k + 1 >= plen || prog[k] != 59 || prog[k + 1] != 112) { ok = false; break; }
            k += 2, ++rounds;
            if (k < plen && prog[k] == 25) { ++k; continue; }
            break;
        }
        if (ok && k == plen - 1 && prog[k] == 56 && uint32_t(rounds) <= m.h_len) {
            mode = CTX_HASHCHAIN;
            n_hash = rounds;
        }
    }
    if (mode == CTX_VM) return;
    m.is_chain = true, m.n_isse = n_isse, m.has_mix2 = mix2, m.ctx_mode = mode, m.n_hash = n_hash;
}

void build_block_prefix(Model &m) {                 // compressor.v:62-75, :150-181
    static const uint8_t loc[16] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3,
                                    0x8c, 0xb2, 0x28, 0xb0, 0xd3, 0x7a, 0x50, 0x51};
    std::vector<uint8_t> &o = m.block_prefix;
    o.assign(loc, loc + 16);
    const int len = int(m.header.size());
    o.push_back((len >= 5 && m.header[4] != 0) ? 1 : 2);
    o.push_back(1);
    const int hsize = (m.cend + 1) + (m.hend - m.hbegin + 1);
    o.push_back(uint8_t(hsize & 0xFF)), o.push_back(uint8_t((hsize >> 8) & 0xFF));
    for (int i = 0; i <= m.cend && i < len; ++i) o.push_back(m.header[size_t(i)]);
    for (int i = m.hbegin; i <= m.hend && i < len; ++i) o.push_back(m.header[size_t(i)]);
}

}  // namespace

int model_from_level_layout(const uint8_t *hdr, int len, Model &m) {
    if (len < 0 || len > 65535 || (len > 0 && !hdr)) return ZPAQGPU_E_ARG;
    m = Model{};
    m.header.assign(hdr, hdr + len);
    if (len >= 5) {                                 // compressor.v:96-140
        const int n = hdr[4];
        int pos = 5;
        for (int i = 0; i < n && pos < len; ++i) {
            const int type = hdr[pos];
            if (type >= 10) break;
            pos += kCompSize[type];
        }
        m.cend = pos;
        if (pos < len && hdr[pos] == 0) ++pos;
        m.hbegin = pos;
        while (pos < len) {
            const uint8_t op = hdr[pos];
            if (op == 0) break;
            ++pos;
            if ((op & 7) == 7) pos += (op == 63) ? 2 : 1;
        }
        m.hend = pos;
    } else {
        m.cend = m.hbegin = m.hend = len;
    }
    const int rc = build_layout(m, false);
    if (rc != ZPAQGPU_OK) return rc;
    detect_shape(m);
    if (m.is_chain) build_layout(m, true);
    build_block_prefix(m);
    return ZPAQGPU_OK;
}

int model_from_archive(const uint8_t *p, uint64_t avail, Model &m, uint64_t *consumed) {
    m = Model{};
    uint64_t at = 0;
    auto get = [&]() -> int { return at < avail ? int(p[at++]) : -1; };
    const int level = get();                        // decompressor.v:257-275
    if (level != 1 && level != 2) return ZPAQGPU_E_FORMAT;
    if (get() != 1) return ZPAQGPU_E_FORMAT;
    const int lo = get(), hi = get();
    if (lo < 0 || hi < 0) return ZPAQGPU_E_FORMAT;
    const int hsize = lo + hi * 256;
    std::vector<uint8_t> &h = m.header;
    for (int i = 0; i < 5; ++i) {                   // decompressor.v:282-288
        const int b = get();
        if (b < 0) return ZPAQGPU_E_FORMAT;
        h.push_back(uint8_t(b));
    }
    const int n = h[4];
    for (int i = 0; i < n; ++i) {                   // decompressor.v:291-306
        const int type = get();
        if (type < 0 || type >= 10) return ZPAQGPU_E_FORMAT;
        h.push_back(uint8_t(type));
        for (int j = 1; j < kCompSize[type]; ++j) {
            const int b = get();
            if (b < 0) return ZPAQGPU_E_FORMAT;
            h.push_back(uint8_t(b));
        }
    }
    if (get() != 0) return ZPAQGPU_E_FORMAT;        // decompressor.v:309-314
    h.push_back(0);
    m.cend = int(h.size()) - 1;
    m.hbegin = int(h.size());
    const int hcomp_len = hsize - int(h.size());    // decompressor.v:322-334
    for (int i = 0; i < hcomp_len; ++i) {
        const int b = get();
        if (b < 0) return ZPAQGPU_E_FORMAT;
        h.push_back(uint8_t(b));
    }
    m.hend = int(h.size()) - 1;
    if (consumed) *consumed = at;
    const int rc = build_layout(m, false);
    if (rc != ZPAQGPU_OK) return rc;
    detect_shape(m);
    if (m.is_chain) build_layout(m, true);
    return ZPAQGPU_OK;
}

}  // namespace zg

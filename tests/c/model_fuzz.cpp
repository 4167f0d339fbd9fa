// Untrusted bytes into the host-side header parser and layout builder (zpaq-v_b200/csrc/model.cpp:
// model_from_archive is what every block header of an archive goes through before any kernel runs).  Built by
// tests/test_model_fuzz.py with -fsanitize=undefined in trap mode and _GLIBCXX_ASSERTIONS (an out-of-range
// vector index aborts): valid headers of all levels, every truncation of them, single-byte mutations at every
// position, and random component tables.  The parser may accept or reject; it may not crash, and what it accepts
// must be internally consistent.
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/zpaqgpu.h"
#include "model.h"

using namespace zg;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state >> 12, rng_state ^= rng_state << 25, rng_state ^= rng_state >> 27;
    return uint32_t((rng_state * 0x2545F4914F6CDD1Dull) >> 32);
}

static long accepted = 0, rejected = 0;

static void check_model(const Model &m) {
    assert(m.n == int(m.comps.size()) || m.comps.empty() || m.n >= 0);
    assert(m.cend >= 0 && m.hbegin == m.cend + 1 && m.hend >= m.cend && m.hend < int(m.header.size()) + 1);
    for (const CompDesc &c : m.comps) {
        assert(c.type >= 0 && c.type <= 9);
        assert(c.cm_off + 4ull * c.cm_len <= m.ws_bytes);
        assert(c.ht_off + uint64_t(c.ht_len) <= m.ws_bytes);
        assert(c.a16_off + 2ull * c.a16_len <= m.ws_bytes);
    }
    for (const FillRegion &f : m.fills) {
        assert(f.off % 4 == 0 && f.off + 4 * f.n_words <= m.ws_bytes);
        assert(f.period >= 1 && uint64_t(f.img_off) + f.period <= m.image.size());
    }
    if (m.is_chain && m.ws_bytes_paged) {
        assert(m.comps_paged.size() == m.comps.size());
        for (const FillRegion &f : m.fills_paged) assert(f.off + 4 * f.n_words <= m.ws_bytes_paged);
    }
}

static void feed(const std::vector<uint8_t> &bytes, uint64_t avail) {
    Model m;
    uint64_t used = 0;
    const int rc = model_from_archive(bytes.data(), avail, m, &used);
    if (rc == ZPAQGPU_OK) {
        assert(used <= avail);
        check_model(m);
        ++accepted;
    } else {
        assert(rc == ZPAQGPU_E_FORMAT || rc == ZPAQGPU_E_UNSUPPORTED || rc == ZPAQGPU_E_NOMEM || rc == ZPAQGPU_E_ARG);
        ++rejected;
    }
}

int main() {
    static const int comp_size[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};   // types.v:74-85
    for (int level = 0; level <= 5; ++level) {
        const std::vector<uint8_t> h = level_header(level);
        Model m;
        assert(model_from_level_layout(h.data(), int(h.size()), m) == ZPAQGPU_OK);
        check_model(m);
        // what start_block writes after the locator: lvl typ hsize COMP HCOMP (compressor.v:157-181)
        assert(m.block_prefix.size() > 16);
        std::vector<uint8_t> tail(m.block_prefix.begin() + 16, m.block_prefix.end());
        tail.resize(tail.size() + 8, 0xEE);
        const long before = accepted;
        feed(tail, tail.size());
        assert(accepted == before + 1);   // the reference's own header is accepted
        for (uint64_t cut = 0; cut < tail.size(); ++cut) feed(tail, cut);          // every truncation
        for (size_t at = 0; at < tail.size(); ++at)                                 // every byte, a few values
            for (int v : {0, 1, 2, 3, 9, 10, 31, 32, 63, 127, 128, 254, 255}) {
                std::vector<uint8_t> t = tail;
                t[at] = uint8_t(v);
                feed(t, t.size());
            }
    }
    // random component tables: valid types with random parameters, and plain noise
    for (int it = 0; it < 200000; ++it) {
        std::vector<uint8_t> comp = {uint8_t(rnd() % 24), uint8_t(rnd() % 24), 0, 0};
        const int n = int(rnd() % 12);
        comp.push_back(uint8_t(n));
        for (int i = 0; i < n; ++i) {
            const int type = 1 + int(rnd() % 9);
            comp.push_back(uint8_t(type));
            for (int j = 1; j < comp_size[type]; ++j)
                comp.push_back(uint8_t((rnd() & 3) ? rnd() % (i + 2 > 40 ? 40 : 34) : rnd()));
        }
        comp.push_back(0);
        const int hlen = int(rnd() % 20);
        std::vector<uint8_t> hcomp;
        for (int i = 0; i < hlen; ++i) hcomp.push_back(uint8_t(rnd()));
        hcomp.push_back(0);
        const size_t hsize = comp.size() + hcomp.size();
        std::vector<uint8_t> t = {uint8_t(n ? 1 : 2), 1, uint8_t(hsize & 255), uint8_t(hsize >> 8)};
        t.insert(t.end(), comp.begin(), comp.end());
        t.insert(t.end(), hcomp.begin(), hcomp.end());
        if (it % 7 == 0)
            for (auto &b : t) b = (rnd() & 15) ? b : uint8_t(rnd());
        feed(t, t.size());
    }
    printf("accepted %ld rejected %ld\nok\n", accepted, rejected);
    return 0;
}

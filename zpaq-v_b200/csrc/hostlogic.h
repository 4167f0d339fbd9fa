// hostlogic.h -- the pure host-side decisions of libzpaqgpu: how a batch is split over devices, where an
// archive may be cut, how many blocks a wave and a CTA take.  No CUDA in here, so the CPU test suite compiles
// and checks these with g++ alone (tests/test_hostlogic.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace zg {

// world+1 boundaries of contiguous unit ranges balanced by bytes ("disjoint block ranges to each
// GPU ... balanced by input bytes"): range g ends at the first unit whose running total reaches
// total * (g+1) / world; with no bytes at all the units themselves are split evenly.
inline std::vector<int> split_by_bytes(const uint64_t *off, int n, int world) {
    std::vector<int> b(size_t(world) + 1, n);
    b[0] = 0;
    const uint64_t base = off[0], total = off[n] - base;
    int k = 1;
    for (int i = 0; i < n && k < world; ++i) {
        const uint64_t acc = off[i + 1] - base;
        while (k < world && (total == 0 ? i + 1 >= (n * k + world - 1) / world
                                        : double(acc) >= double(total) * k / world))
            b[size_t(k++)] = i + 1;
    }
    return b;
}

// offset of the first 13-byte locator + "zPQ" at or after `from` (Decompresser.find_block looks for these 16
// bytes, decompressor.v:227-254), or len
inline uint64_t next_locator(const uint8_t *arc, uint64_t len, uint64_t from) {
    static const uint8_t tag[16] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3,
                                    0x8c, 0xb2, 0x28, 0xb0, 0xd3, 'z', 'P', 'Q'};
    uint64_t p = from;
    while (p + 16 <= len) {
        const void *hit = std::memchr(arc + p, tag[0], size_t(len - 15 - p));
        if (!hit) return len;
        p = uint64_t(static_cast<const uint8_t *>(hit) - arc);
        if (std::memcmp(arc + p, tag, 16) == 0) return p;
        ++p;
    }
    return len;
}

// Blocks per dense wave.  `mem_slots` blocks fit the table memory, the codec kernel holds `round_cap` blocks at
// once (SMs x blocks per CTA; 0 = do not care).  A wave larger than round_cap runs its CTAs in two rounds, the
// second nearly empty, so a batch of several waves is cut into as few rounds as its size needs, all equal.
inline uint64_t wave_slots(uint64_t n_blocks, uint64_t mem_slots, uint64_t round_cap) {
    uint64_t slots = std::min(mem_slots, n_blocks);
    if (round_cap > 0 && n_blocks > slots && slots > 0) {
        const uint64_t per_round = std::min(slots, round_cap);
        const uint64_t rounds = (n_blocks + per_round - 1) / per_round;
        slots = (n_blocks + rounds - 1) / rounds;
    }
    return slots;
}

// Blocks per CTA of a launch of n blocks when a CTA holds at most `most` and one CTA fits an SM: as many as
// spread the launch over all SMs; when that is more than a CTA holds, the CTAs run in rounds and the rounds
// are made equal instead of a full one followed by a nearly empty one.
inline int blocks_per_cta(int n, int most, int sms) {
    sms = std::max(1, sms);
    const int per_sm = (n + sms - 1) / sms;
    if (per_sm <= most) return std::max(1, per_sm);
    const int rounds = (n + sms * most - 1) / (sms * most);
    return std::max(1, std::min(most, (n + rounds * sms - 1) / (rounds * sms)));
}

// itos_pad / make_jidac_filename (jidac.v:38-49), the block comment "<usize> jDC\x01" (jidac.v:69, :96) and
// little-endian fields of the index blocks
inline std::string pad_num(long long n, int width) {
    std::string s = std::to_string(n);
    while (int(s.size()) < width) s = "0" + s;
    return s;
}
inline std::string jidac_name(long long date, char type, uint32_t num) {
    return "jDC" + pad_num(date, 14) + std::string(1, type) + pad_num(num, 10);
}
inline std::string jidac_comment(uint64_t usize) { return std::to_string(usize) + " jDC\x01"; }
inline void put_le(std::vector<uint8_t> &v, uint64_t x, int bytes) {
    for (int i = 0; i < bytes; ++i) v.push_back(uint8_t(x >> (8 * i)));
}

}  // namespace zg

// CPU check of zpaq-v_b200/csrc/hostlogic.h (g++ only, no CUDA): prints one line per case, "ok" at the end.
#include <cassert>
#include <cstdio>
#include <numeric>
#include <vector>

#include "hostlogic.h"

using namespace zg;

static std::vector<uint64_t> offsets(const std::vector<uint64_t> &sizes) {
    std::vector<uint64_t> off(sizes.size() + 1, 7);  // a base other than zero
    for (size_t i = 0; i < sizes.size(); ++i) off[i + 1] = off[i] + sizes[i];
    return off;
}

int main() {
    // ---- split_by_bytes: contiguous, covering, balanced by bytes ----
    for (int world : {1, 2, 3, 8}) {
        for (const std::vector<uint64_t> &sizes :
             {std::vector<uint64_t>(16, 1000), std::vector<uint64_t>{5, 9000, 3, 3, 3, 9000, 1, 1},
              std::vector<uint64_t>{0, 0, 0, 0, 0}, std::vector<uint64_t>{42}, std::vector<uint64_t>{}}) {
            const auto off = offsets(sizes);
            const int n = int(sizes.size());
            const auto b = split_by_bytes(off.data(), n, world);
            assert(int(b.size()) == world + 1 && b.front() == 0 && b.back() == n);
            for (int g = 0; g < world; ++g) assert(b[g] <= b[g + 1]);
            const uint64_t total = off[n] - off[0];
            if (total) {  // no range exceeds its fair share by more than one unit's bytes
                uint64_t largest = 0;
                for (uint64_t s : sizes) largest = std::max(largest, s);
                for (int g = 0; g < world; ++g) {
                    const uint64_t got = off[b[g + 1]] - off[b[g]];
                    assert(got <= total / world + largest + 1);
                }
            } else if (n) {   // no bytes: units split evenly
                for (int g = 0; g < world; ++g) assert(b[g + 1] - b[g] <= (n + world - 1) / world);
            }
        }
    }
    {   // equal blocks, two devices: halves
        const auto off = offsets(std::vector<uint64_t>(1024, 1 << 20));
        const auto b = split_by_bytes(off.data(), 1024, 2);
        assert(b[1] == 512);
    }
    // ---- next_locator ----
    {
        const uint8_t tag[16] = {0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3, 'z', 'P', 'Q'};
        std::vector<uint8_t> arc(5000, 0x37);              // many false first bytes
        std::memcpy(arc.data() + 100, tag, 16);
        std::memcpy(arc.data() + 3000, tag, 16);
        std::memcpy(arc.data() + 4990, tag, 10);           // cut off at the end: not a locator
        assert(next_locator(arc.data(), arc.size(), 0) == 100);
        assert(next_locator(arc.data(), arc.size(), 100) == 100);
        assert(next_locator(arc.data(), arc.size(), 101) == 3000);
        assert(next_locator(arc.data(), arc.size(), 3001) == arc.size());
        assert(next_locator(arc.data(), 10, 0) == 10);
        assert(next_locator(arc.data(), 0, 0) == 0);
    }
    // ---- wave_slots ----
    assert(wave_slots(8192, 1127, 1036) == 1024);          // cfg 3 encoder: 8 equal rounds of 147 CTAs
    assert(wave_slots(8192, 1250, 1184) == 1171);          // cfg 3 decoder: 7 waves, 8 warps per CTA, one round each
    assert(wave_slots(8, 3, 1036) == 3);                   // memory is the limit
    assert(wave_slots(1024, 5000, 1036) == 1024);          // one wave: nothing to cut
    assert(wave_slots(1100, 1127, 0) == 1100);
    assert(wave_slots(2100, 1127, 1036) == 700);           // three rounds of 700 rather than 1036 + 1036 + 28
    for (uint64_t n = 1; n < 5000; n += 37)
        for (uint64_t mem : {1ull, 56ull, 1127ull}) {
            const uint64_t s = wave_slots(n, mem, 1036);
            assert(s >= 1 && s <= std::min(mem, n));
        }
    // ---- blocks_per_cta ----
    assert(blocks_per_cta(1024, 7, 148) == 7);             // cfg 2: 147 CTAs, one round
    assert(blocks_per_cta(1024, 6, 148) == 4);             // -m5 encoder: two rounds of 128 CTAs instead of 148 + 23
    assert(blocks_per_cta(37, 7, 148) == 1);
    assert(blocks_per_cta(10000, 7, 148) == 7);
    assert(blocks_per_cta(0, 7, 148) == 1);
    for (int n = 1; n < 3000; n += 13) {
        const int b = blocks_per_cta(n, 6, 148);
        assert(b >= 1 && b <= 6);
    }
    std::puts("ok");
    return 0;
}

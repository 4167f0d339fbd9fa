// sha1_lane.h -- what ONE lane of k_sha1_staged does with its byte range (sha1.v:42-146): which 16-byte
// chunks of the range a staging round brings into shared memory, how the 16 big-endian words of a 64-byte
// block come out of the staged window at any byte alignment, and how the padding blocks (0x80, zeros, bit
// length; sha1.v:112-134) are made from the same window.  Plain C++ behind ZG_HD, so that the CPU test suite
// runs the identical code under a host emulation of the warp (tests/c/sha1_lane_test.cpp).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ZG_HD __host__ __device__ __forceinline__
#else
#define ZG_HD inline
#endif

namespace zg {
namespace sha1lane {

constexpr uint32_t kBlocksPerRound = 4;                        // 64-byte SHA-1 blocks per staging round
constexpr uint32_t kRoundBytes = 64 * kBlocksPerRound;         // 256
constexpr uint32_t kChunksPerRow = kRoundBytes / 16 + 1;       // 17: a block may start up to 15 bytes into a chunk
constexpr uint32_t kRowBytes = 16 * kChunksPerRow;             // 272
constexpr uint32_t kRowWords = kRowBytes / 4;                  // 68

ZG_HD uint32_t rol(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(x, x, n);
#else
    return (x << n) | (x >> (32 - n));
#endif
}

// Bytes sh..sh+3 of the little-endian pair (lo, hi) as one big-endian word (sh = 0..3).
ZG_HD uint32_t be_word(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(lo, hi, (sh << 12) | ((sh + 1u) << 8) | ((sh + 2u) << 4) | (sh + 3u));
#else
    const uint64_t pair = (uint64_t(hi) << 32) | lo;
    const uint32_t v = uint32_t(pair >> (8 * sh));
    return (v << 24) | ((v & 0xFF00u) << 8) | ((v >> 8) & 0xFF00u) | (v >> 24);
#endif
}

ZG_HD void compress(uint32_t st[5], uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 80; ++i) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            wi = rol(w[(i - 3) & 15] ^ w[(i - 8) & 15] ^ w[(i - 14) & 15] ^ w[i & 15], 1);
            w[i & 15] = wi;
        }
        uint32_t f, k;
        if (i < 20) f = (b & c) | (~b & d), k = 0x5A827999u;
        else if (i < 40) f = b ^ c ^ d, k = 0x6ED9EBA1u;
        else if (i < 60) f = (b & c) | (b & d) | (c & d), k = 0x8F1BBCDCu;
        else f = b ^ c ^ d, k = 0xCA62C1D6u;
        const uint32_t t = rol(a, 5) + f + e + k + wi;
        e = d, d = c, c = rol(b, 30), b = a, a = t;
    }
    st[0] += a, st[1] += b, st[2] += c, st[3] += d, st[4] += e;
}

// One byte range as its lane sees it.  Addresses are plain integers so that host and device agree.
struct Lane {
    uint64_t s, end;        // first byte, one past the last
    uint64_t a;             // s rounded down to 16: the staged window of round r starts at a + 256 r
    uint32_t shift;         // s - a
    uint32_t full, total;   // whole 64-byte blocks of data; blocks including the padding (0 = no job)
    uint32_t st[5];
};

ZG_HD void lane_begin(Lane &L, uint64_t addr, uint64_t len, bool has_job) {
    L.s = addr, L.end = addr + len;
    L.a = addr & ~uint64_t(15);
    L.shift = uint32_t(addr - L.a);
    L.full = uint32_t(len / 64);
    const uint32_t rem = uint32_t(len & 63);
    L.total = has_job ? L.full + (rem >= 56 ? 2u : 1u) : 0u;   // 0x80 and the 8 length bytes must fit
    L.st[0] = 0x67452301u, L.st[1] = 0xEFCDAB89u, L.st[2] = 0x98BADCFEu, L.st[3] = 0x10325476u, L.st[4] = 0xC3D2E1F0u;
}

ZG_HD uint32_t lane_rounds(const Lane &L) { return (L.total + kBlocksPerRound - 1) / kBlocksPerRound; }

// What chunk c of round r of this lane's window needs: kWhole = all 16 bytes belong to the range (one
// 16-byte copy), kPart = only bytes [lo, hi) of it do (the ends of an unaligned range: copied byte by byte,
// nothing outside the range is ever read), kNone = nothing.
enum ChunkKind { kNone = 0, kWhole = 1, kPart = 2 };
ZG_HD ChunkKind chunk_plan(uint64_t a, uint64_t s, uint64_t end, uint32_t r, uint32_t c, uint64_t &src, uint32_t &lo,
                           uint32_t &hi) {
    src = a + uint64_t(r) * kRoundBytes + 16u * c;
    if (src >= end || src + 16 <= s) return kNone;
    if (src >= s && src + 16 <= end) return kWhole;
    lo = src < s ? uint32_t(s - src) : 0u;
    hi = src + 16 > end ? uint32_t(end - src) : 16u;
    return kPart;
}

// The 16 message words of block 4 r + k from the lane's staged row (kRowWords words, little-endian as loaded).
// Data blocks come straight from the window; the blocks after them are the padding of sha1.v:112-134, made
// from whatever data bytes the window still holds (bytes past `end` in the window are never looked at).
ZG_HD void block_words(const Lane &L, const uint32_t *row, uint32_t r, uint32_t k, uint32_t w[16]) {
    const uint32_t b = r * kBlocksPerRound + k;
    const uint32_t o = L.shift + 64u * k;
    const uint32_t *p = row + (o >> 2);
    const uint32_t sh = o & 3u;
    uint32_t x[17];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 17; ++i) x[i] = p[i];   // p[16] is inside the row: (15 + 192) / 4 + 16 = 67 < 68
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; ++i) w[i] = be_word(x[i], x[i + 1], sh);
    if (b >= L.full) {
        const uint64_t len = L.end - L.s;
        const uint32_t valid = b == L.full ? uint32_t(len & 63) : 0u;   // data bytes at the front of this block
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 16; ++i) {
            const int nv = int(valid) - 4 * i;      // data bytes in word i
            uint32_t v = nv >= 4 ? w[i] : (nv <= 0 ? 0u : (w[i] & (0xFFFFFFFFu << (32 - 8 * nv))));
            if (b == L.full && int(valid >> 2) == i) v |= 0x80u << (24 - 8 * (valid & 3u));
            w[i] = v;
        }
        if (b + 1 == L.total) {
            const uint64_t bits = len * 8;
            w[14] = uint32_t(bits >> 32), w[15] = uint32_t(bits);
        }
    }
}

ZG_HD void digest_bytes(const Lane &L, uint8_t *out) {
    for (int i = 0; i < 5; ++i) {
        out[i * 4] = uint8_t(L.st[i] >> 24), out[i * 4 + 1] = uint8_t(L.st[i] >> 16);
        out[i * 4 + 2] = uint8_t(L.st[i] >> 8), out[i * 4 + 3] = uint8_t(L.st[i]);
    }
}

}  // namespace sha1lane
}  // namespace zg

// CPU emulation of one warp of k_sha1_staged (zpaq-v_b200/csrc/kernels_aux.cu) over the lane logic of
// zpaq-v_b200/csrc/sha1_lane.h (g++ only, no CUDA).  The 32 lanes run one after the other inside each phase
// the kernel separates with __syncwarp; cp.async is a 16-byte memcpy that ASSERTS its source lies inside the
// byte range of its job, as do the byte copies of the partial chunks; the two stage buffers start out as
// garbage and are never cleared, like shared memory.
//
// stdin:  u64 n_jobs, u64 data_len, u64 base_shift, n_jobs x {u64 off, u64 len}, data bytes
// stdout: one hex digest per job
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sha1_lane.h"

using namespace zg::sha1lane;

struct Job {
    uint64_t off, len;
};

static void run_warp(const uint8_t *base, const Job *jobs, int n_jobs, int first, uint8_t *digests) {
    static uint8_t buf[2][32 * kRowBytes];
    memset(buf, 0xA5, sizeof buf);
    Lane L[32];
    uint64_t desc[32][3];
    uint32_t n_rounds = 0;
    for (int lane = 0; lane < 32; ++lane) {
        const int j = first + lane;
        const bool has = j < n_jobs;
        lane_begin(L[lane], reinterpret_cast<uint64_t>(base) + (has ? jobs[j].off : 0), has ? jobs[j].len : 0, has);
        desc[lane][0] = L[lane].a, desc[lane][1] = L[lane].s, desc[lane][2] = L[lane].end;
        if (lane_rounds(L[lane]) > n_rounds) n_rounds = lane_rounds(L[lane]);
    }
    auto issue = [&](uint32_t r) {
        for (int lane = 0; lane < 32; ++lane) {
            for (uint32_t i = 0; i < kChunksPerRow; ++i) {
                const uint32_t q = uint32_t(lane) + 32u * i;
                const uint32_t row = q / kChunksPerRow, c = q - row * kChunksPerRow;
                uint64_t src;
                uint32_t lo = 0, hi = 0;
                const ChunkKind kind = chunk_plan(desc[row][0], desc[row][1], desc[row][2], r, c, src, lo, hi);
                uint8_t *dst = buf[r & 1] + row * kRowBytes + 16u * c;
                if (kind == kWhole) {
                    assert(src >= desc[row][1] && src + 16 <= desc[row][2] && (src & 15) == 0);
                    memcpy(dst, reinterpret_cast<const void *>(src), 16);
                } else if (kind == kPart) {
                    for (uint32_t t = lo; t < hi; ++t) {
                        assert(src + t >= desc[row][1] && src + t < desc[row][2]);
                        dst[t] = *reinterpret_cast<const uint8_t *>(src + t);
                    }
                }
            }
        }
    };
    if (n_rounds) issue(0);
    for (uint32_t r = 0; r < n_rounds; ++r) {
        if (r + 1 < n_rounds) issue(r + 1);
        for (int lane = 0; lane < 32; ++lane) {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(buf[r & 1] + lane * kRowBytes);
            for (uint32_t k = 0; k < kBlocksPerRound; ++k) {
                if (r * kBlocksPerRound + k < L[lane].total) {
                    uint32_t w[16];
                    block_words(L[lane], row, r, k, w);
                    compress(L[lane].st, w);
                }
            }
        }
    }
    for (int lane = 0; lane < 32; ++lane)
        if (first + lane < n_jobs) digest_bytes(L[lane], digests + size_t(first + lane) * 20);
}

int main() {
    uint64_t head[3];
    if (fread(head, 8, 3, stdin) != 3) return 2;
    const uint64_t n_jobs = head[0], data_len = head[1], base_shift = head[2] & 63;
    std::vector<Job> jobs(n_jobs);
    if (n_jobs && fread(jobs.data(), sizeof(Job), n_jobs, stdin) != n_jobs) return 2;
    // the data sits at an address with the requested misalignment, exactly as long as it is
    uint8_t *raw = static_cast<uint8_t *>(aligned_alloc(64, size_t((data_len + base_shift + 63) / 64 * 64 + 64)));
    uint8_t *base = raw + base_shift;
    if (data_len && fread(base, 1, data_len, stdin) != data_len) return 2;
    for (const Job &j : jobs) assert(j.off + j.len <= data_len);
    std::vector<uint8_t> dig(n_jobs * 20 + 1);
    for (uint64_t first = 0; first < n_jobs; first += 32) run_warp(base, jobs.data(), int(n_jobs), int(first), dig.data());
    for (uint64_t j = 0; j < n_jobs; ++j) {
        for (int i = 0; i < 20; ++i) printf("%02x", dig[j * 20 + i]);
        printf("\n");
    }
    free(raw);
    return 0;
}

// multi.cu -- one box, several B200s behind ONE C handle (SURVEY 8(b) `zpaqgpu_init(ctx, devices,
// n_devices)`, 8(e)).  ZPAQ blocks are independent (compressor.v:84-187 re-creates every piece of model
// state in start_block), so the split is by contiguous block ranges balanced by input bytes, one host
// thread + context + stream per device, no collective; results land in the caller's buffer in block
// order.  Two phases per call: every device stages its range (upload + kernels, sizes become known),
// then every device copies its bytes to its final place in the caller's buffer.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ctx.h"
#include "hostlogic.h"

using namespace zg;

struct zpaqgpu_multi {
    std::vector<zpaqgpu_ctx *> ctx;
    std::vector<int> devices;
    std::string err;
    // what the last call gave every device (zpaqgpu_multi_last_stats)
    std::vector<int> first_unit, n_units;
    std::vector<zpaqgpu_stats> stats;
    std::vector<float> stage_ms, fetch_ms;
    int fallback_single = 0;  // last decompress call was repeated on one device (see below)
};

namespace {

template <class F>
void on_every_device(int n, F &&f) {
    std::vector<std::thread> th;
    for (int g = 1; g < n; ++g) th.emplace_back([&f, g] { f(g); });
    f(0);
    for (auto &t : th) t.join();
}

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int first_error(zpaqgpu_multi *m, const std::vector<int> &rcs) {
    for (size_t g = 0; g < rcs.size(); ++g)
        if (rcs[g] != ZPAQGPU_OK) {
            m->err = "device " + std::to_string(m->devices[g]) + ": " + m->ctx[g]->err;
            return rcs[g];
        }
    return ZPAQGPU_OK;
}

}  // namespace

extern "C" {

int zpaqgpu_multi_init(zpaqgpu_multi **out, const int *devices, int n_devices) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    if (!out) return ZPAQGPU_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return ZPAQGPU_E_NODEVICE;
    }
    std::vector<int> dev;
    if (!devices || n_devices <= 0)
        for (int d = 0; d < count; ++d) dev.push_back(d);
    else
        dev.assign(devices, devices + n_devices);
    for (int d : dev)
        if (d < 0 || d >= count) return ZPAQGPU_E_ARG;
    zpaqgpu_multi *m = new zpaqgpu_multi();
    m->devices = dev;
    for (int d : dev) {
        zpaqgpu_ctx *c = nullptr;
        const int rc = zpaqgpu_init(&c, d);
        if (rc != ZPAQGPU_OK) {
            for (zpaqgpu_ctx *x : m->ctx) zpaqgpu_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    const size_t n = dev.size();
    m->first_unit.assign(n, 0), m->n_units.assign(n, 0), m->stats.assign(n, zpaqgpu_stats{});
    m->stage_ms.assign(n, 0.f), m->fetch_ms.assign(n, 0.f);
    *out = m;
    return ZPAQGPU_OK;
    });
}

void zpaqgpu_multi_destroy(zpaqgpu_multi *m) {
    if (!m) return;
    for (zpaqgpu_ctx *c : m->ctx) zpaqgpu_destroy(c);
    delete m;
}

int zpaqgpu_multi_device_count(const zpaqgpu_multi *m) { return m ? int(m->ctx.size()) : 0; }

zpaqgpu_ctx *zpaqgpu_multi_ctx(zpaqgpu_multi *m, int k) {
    return (m && k >= 0 && k < int(m->ctx.size())) ? m->ctx[size_t(k)] : nullptr;
}

const char *zpaqgpu_multi_last_error(const zpaqgpu_multi *m) { return m ? m->err.c_str() : ""; }

int zpaqgpu_multi_last_stats(const zpaqgpu_multi *m, int k, zpaqgpu_multi_stats *out) {
    if (!m || !out || k < 0 || k >= int(m->ctx.size())) return ZPAQGPU_E_ARG;
    out->device = m->devices[size_t(k)];
    out->first_unit = m->first_unit[size_t(k)], out->n_units = m->n_units[size_t(k)];
    out->fallback_single = m->fallback_single;
    out->stage_ms = m->stage_ms[size_t(k)], out->fetch_ms = m->fetch_ms[size_t(k)];
    out->stats = m->stats[size_t(k)];
    return ZPAQGPU_OK;
}

int zpaqgpu_multi_compress_blocks(zpaqgpu_multi *m, int level, const uint8_t *in, const uint64_t *in_off, int n_blocks,
                                  const char *const *names, const char *const *comments, uint8_t *out,
                                  uint64_t out_cap, uint64_t *out_off, uint64_t *out_need) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    if (!m || n_blocks < 0 || (n_blocks > 0 && (!in_off || !out_off))) return ZPAQGPU_E_ARG;
    if (out_need) *out_need = 0;
    if (n_blocks == 0) {
        if (out_off) out_off[0] = 0;
        return ZPAQGPU_OK;
    }
    for (int b = 0; b < n_blocks; ++b)
        if (in_off[b + 1] < in_off[b]) return ZPAQGPU_E_ARG;
    const std::vector<uint8_t> h = level_header(level);
    Model model;
    int rc = model_from_level_layout(h.data(), int(h.size()), model);
    if (rc) return m->err = model.error, rc;
    const int G = int(m->ctx.size());
    const std::vector<int> cut = split_by_bytes(in_off, n_blocks, G);
    std::vector<int> rcs(size_t(G), ZPAQGPU_OK);
    std::vector<u64> totals(size_t(G), 0);
    m->fallback_single = 0;
    // phase 1: upload + kernels on every device; the archive of a range stays on its device
    on_every_device(G, [&](int g) {
        const int lo = cut[size_t(g)], n = cut[size_t(g) + 1] - lo;
        m->first_unit[size_t(g)] = lo, m->n_units[size_t(g)] = n;
        m->stats[size_t(g)] = zpaqgpu_stats{};
        m->stage_ms[size_t(g)] = m->fetch_ms[size_t(g)] = 0.f;
        if (n == 0) return;
        const double t0 = now_ms();
        rcs[size_t(g)] = zg::guarded<int>(m->ctx[size_t(g)], [&]() -> int {
            return compress_stage(m->ctx[size_t(g)], model, in, in_off + lo, n, names ? names + lo : nullptr,
                                  comments ? comments + lo : nullptr, &totals[size_t(g)]);
        });
        m->stage_ms[size_t(g)] = float(now_ms() - t0);
    });
    if ((rc = first_error(m, rcs))) return rc;
    std::vector<u64> base(size_t(G) + 1, 0);
    for (int g = 0; g < G; ++g) base[size_t(g) + 1] = base[size_t(g)] + totals[size_t(g)];
    if (out_need) *out_need = base[size_t(G)];
    if (base[size_t(G)] > out_cap) return ZPAQGPU_E_NOSPACE;
    if (base[size_t(G)] && !out) return ZPAQGPU_E_ARG;
    // phase 2: every device copies its bytes to their final place; offsets are shifted by the range's base.
    // Range g writes out_off[lo..hi]; entry hi is written again (same value) by range g+1's first entry.
    on_every_device(G, [&](int g) {
        const int lo = cut[size_t(g)], n = cut[size_t(g) + 1] - lo;
        if (n == 0) return;
        const double t0 = now_ms();
        std::vector<u64> off(size_t(n) + 1);
        rcs[size_t(g)] = zg::guarded<int>(m->ctx[size_t(g)], [&]() -> int {
            return compress_fetch(m->ctx[size_t(g)], n, totals[size_t(g)], out + base[size_t(g)], off.data(),
                                  base[size_t(g)]);
        });
        // interior entries only: the boundary entries are set below from `base`, so that no two threads
        // write the same word
        for (int b = 1; b < n; ++b) out_off[lo + b] = off[size_t(b)];
        m->stats[size_t(g)] = m->ctx[size_t(g)]->stats;
        m->fetch_ms[size_t(g)] = float(now_ms() - t0);
    });
    for (int g = 0; g <= G; ++g) out_off[cut[size_t(g)]] = base[size_t(std::min(g, G))];
    return first_error(m, rcs);
    });
}

int zpaqgpu_multi_decompress_archive(zpaqgpu_multi *m, const uint8_t *arc, uint64_t len, uint8_t *out, uint64_t out_cap,
                                     uint64_t *out_need, zpaqgpu_segment *segs, int segs_cap, int *n_segs) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    if (!m || (len && !arc)) return ZPAQGPU_E_ARG;
    if (out_need) *out_need = 0;
    if (n_segs) *n_segs = 0;
    if (len == 0) return ZPAQGPU_OK;
    const int G = int(m->ctx.size());
    m->fallback_single = 0;
    // Byte ranges: range g starts at the first locator at or after len*g/G (found by a host memchr scan of
    // that neighbourhood only; every device still runs the reference's rolling-hash scan, k_find_blocks,
    // over its own range).  A range that finds no locator is empty.
    std::vector<u64> cut(size_t(G) + 1, len);
    cut[0] = 0;
    for (int g = 1; g < G; ++g) cut[size_t(g)] = std::max(cut[size_t(g) - 1], next_locator(arc, len, len / u64(G) * u64(g)));
    struct Part {
        std::vector<DecodedSeg> list;
        const u8 *d_plain = nullptr;
        int status = ZPAQGPU_OK;
        bool stopped = false;
        u64 total = 0;
    };
    std::vector<Part> part(static_cast<size_t>(G));
    std::vector<int> rcs(size_t(G), ZPAQGPU_OK);
    auto stage = [&](int g, u64 a, u64 b) {
        m->first_unit[size_t(g)] = 0, m->n_units[size_t(g)] = 0;
        m->stats[size_t(g)] = zpaqgpu_stats{};
        m->stage_ms[size_t(g)] = m->fetch_ms[size_t(g)] = 0.f;
        if (b <= a) return;
        const double t0 = now_ms();
        Part &p = part[size_t(g)];
        rcs[size_t(g)] = zg::guarded<int>(m->ctx[size_t(g)], [&]() -> int {
            return decode_archive_dev(m->ctx[size_t(g)], arc + a, b - a, p.list, &p.d_plain, &p.status, &p.total);
        });
        p.stopped = m->ctx[size_t(g)]->walk_stopped;
        m->stage_ms[size_t(g)] = float(now_ms() - t0);
    };
    on_every_device(G, [&](int g) { stage(g, cut[size_t(g)], cut[size_t(g) + 1]); });
    int rc = first_error(m, rcs);
    if (rc) return rc;
    // The ranges only compose to what repeated find_block calls over the whole archive give when every
    // range but the last ends cleanly: a block that runs into the next range (the locator that started
    // that range lay INSIDE a block: an archive stored in an archive) or a damaged block in the middle (the
    // reference stops there) shows up as a non-OK status.  Then one device walks the whole archive.
    int last = 0;
    for (int g = 0; g < G; ++g)
        if (cut[size_t(g) + 1] > cut[size_t(g)]) last = g;
    bool clean = true;
    for (int g = 0; g < last; ++g)
        if (cut[size_t(g) + 1] > cut[size_t(g)] && (part[size_t(g)].status != ZPAQGPU_OK || part[size_t(g)].stopped))
            clean = false;
    if (!clean) {
        m->fallback_single = 1;
        for (Part &p : part) p = Part{};
        for (int g = 0; g <= G; ++g) cut[size_t(g)] = g == 0 ? 0 : len;
        stage(0, 0, len);
        if ((rc = first_error(m, rcs))) return rc;
        last = 0;
    }
    u64 total = 0;
    int seg_total = 0;
    for (const Part &p : part) total += p.total, seg_total += int(p.list.size());
    if (out_need) *out_need = total;
    if (n_segs) *n_segs = seg_total;
    if (total > out_cap || (segs && seg_total > segs_cap)) return ZPAQGPU_E_NOSPACE;
    if (total && !out) return ZPAQGPU_E_ARG;
    std::vector<u64> obase(size_t(G) + 1, 0);
    std::vector<int> sbase(size_t(G) + 1, 0), bbase(size_t(G) + 1, 0);
    for (int g = 0; g < G; ++g) {
        const Part &p = part[size_t(g)];
        obase[size_t(g) + 1] = obase[size_t(g)] + p.total;
        sbase[size_t(g) + 1] = sbase[size_t(g)] + int(p.list.size());
        bbase[size_t(g) + 1] = bbase[size_t(g)] + (p.list.empty() ? 0 : p.list.back().seg.block_index + 1);
    }
    on_every_device(G, [&](int g) {
        Part &p = part[size_t(g)];
        m->first_unit[size_t(g)] = bbase[size_t(g)], m->n_units[size_t(g)] = bbase[size_t(g) + 1] - bbase[size_t(g)];
        if (p.list.empty()) return;
        const double t0 = now_ms();
        rcs[size_t(g)] = zg::guarded<int>(m->ctx[size_t(g)], [&]() -> int {
            return plain_fetch(m->ctx[size_t(g)], p.list, p.d_plain, out + obase[size_t(g)]);
        });
        if (segs)
            for (size_t k = 0; k < p.list.size(); ++k) {
                zpaqgpu_segment s = p.list[k].seg;
                s.block_start += cut[size_t(g)], s.block_end += cut[size_t(g)];
                s.name_off += cut[size_t(g)], s.comment_off += cut[size_t(g)];
                s.out_off += obase[size_t(g)];
                s.block_index += bbase[size_t(g)];
                segs[size_t(sbase[size_t(g)]) + k] = s;
            }
        m->stats[size_t(g)] = m->ctx[size_t(g)]->stats;
        m->fetch_ms[size_t(g)] = float(now_ms() - t0);
    });
    if ((rc = first_error(m, rcs))) return rc;
    return part[size_t(last)].status;
    });
}

}  // extern "C"

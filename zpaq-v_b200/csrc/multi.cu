// multi.cu -- one box, several B200s behind ONE C handle (SURVEY 8(b) `zpaqgpu_init(ctx, devices,
// n_devices)`, 8(e)).  ZPAQ blocks are independent (compressor.v:84-187 re-creates every piece of model
// state in start_block), so the split is by contiguous block ranges balanced by input bytes, one host
// thread + context + stream per device, no collective; results land in the caller's buffer in block
// order.  Two phases per call: every device stages its range (upload + kernels, sizes become known),
// then every device copies its bytes to its final place in the caller's buffer.
#include <algorithm>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <unordered_map>
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
    zpaqgpu_ctx *aux = nullptr;  // second context on the first device: index blocks of jidac add, coded while
                                 // that device still holds its d blocks
};

namespace {

// f(g) on one host thread per device (device 0 on the caller's).  An exception in any of them (an allocation
// of host memory that fails) is held until every thread has been joined and then rethrown on the caller's
// thread, where the guarded entry point turns it into a status: no exception leaves a std::thread and no
// joinable thread is ever destroyed.
template <class F>
void on_every_device(int n, F &&f) {
    std::vector<std::exception_ptr> thrown(static_cast<size_t>(std::max(n, 1)));
    auto run = [&f, &thrown](int g) {
        try {
            f(g);
        } catch (...) {
            thrown[size_t(g)] = std::current_exception();
        }
    };
    std::vector<std::thread> th;
    try {
        for (int g = 1; g < n; ++g) th.emplace_back(run, g);
    } catch (...) {   // a thread could not be started: the devices without one are reported, the others joined
        thrown[0] = std::current_exception();
    }
    if (!thrown[0]) run(0);
    for (auto &t : th) t.join();
    for (const std::exception_ptr &e : thrown)
        if (e) std::rethrow_exception(e);
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
    if (m->aux) zpaqgpu_destroy(m->aux);
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

// `jidac add` over all devices of the handle (north_star: "jidac add ... 8 x B200").  The one path with an
// exchange step: a duplicate of a file that another device holds can only be found through the fragment
// digests of every device.
//   1. files are split into contiguous ranges balanced by bytes; every device cuts and hashes its range
//      (zpaqgpu_jidac_fragment: rolling-hash fragmentation + SHA-1 kernels);
//   2. EXCHANGE, on the host: the (SHA-1, length) lists are merged in file order into one table -- ids count
//      first occurrences (jidac.v:153-163), a later equal fragment, on whichever device, takes the id of the
//      first; 24 bytes per fragment cross the host, no payload does;
//   3. every device packs the fragments it stores (first occurrences inside its range) into d blocks of up
//      to block_bytes and codes them; a d block never spans devices, so the block cut depends on the number
//      of devices, the contents of the archive do not (one device: the bytes of zpaqgpu_jidac_add);
//   4. c, h and i blocks (jidac.v:216-295) are written by the first device; every device then copies its d
//      blocks to their place behind the c block.
int zpaqgpu_multi_jidac_add(zpaqgpu_multi *m, const zpaqgpu_jidac_opts *opts, const char *const *names,
                            const uint8_t *in, const uint64_t *in_off, int n_files, uint8_t *out, uint64_t out_cap,
                            uint64_t *out_len, uint64_t *out_need) {
    return zg::guarded<int>(static_cast<zpaqgpu_ctx *>(nullptr), [&]() -> int {
    if (!m || !opts || n_files < 0 || (n_files > 0 && (!in_off || !names))) return ZPAQGPU_E_ARG;
    if (opts->level < 0 || opts->level > 5 || opts->fragment < -1 || opts->fragment > 40) return ZPAQGPU_E_ARG;
    if (out_len) *out_len = 0;
    if (out_need) *out_need = 0;
    const int G = int(m->ctx.size());
    m->fallback_single = 0;
    std::vector<int> cut(size_t(G) + 1, 0);
    if (n_files > 0) cut = split_by_bytes(in_off, n_files, G);
    std::vector<int> rcs(size_t(G), ZPAQGPU_OK);
    int rc;
    // 1. fragments of every range
    std::vector<std::vector<zpaqgpu_fragment>> frs(static_cast<size_t>(G));
    on_every_device(G, [&](int g) {
        const int lo = cut[size_t(g)], n = cut[size_t(g) + 1] - lo;
        m->first_unit[size_t(g)] = lo, m->n_units[size_t(g)] = n;
        m->stats[size_t(g)] = zpaqgpu_stats{};
        m->stage_ms[size_t(g)] = m->fetch_ms[size_t(g)] = 0.f;
        if (n == 0) return;
        const double t0 = now_ms();
        const u64 bytes = in_off[lo + n] - in_off[lo];
        const u64 smallest = opts->fragment < 0 ? ~0ull : (64ull << std::min(opts->fragment, 40));
        int cap = int(std::min<u64>(bytes / std::max<u64>(smallest, 1) + u64(n) + 16, 1u << 30));
        for (int attempt = 0; attempt < 2; ++attempt) {
            frs[size_t(g)].resize(size_t(cap));
            int nf = 0, ns = 0;
            rcs[size_t(g)] = zpaqgpu_jidac_fragment(m->ctx[size_t(g)], in, in_off + lo, n, opts->fragment, 0,
                                                    frs[size_t(g)].data(), cap, &nf, &ns);
            if (rcs[size_t(g)] == ZPAQGPU_E_NOSPACE && attempt == 0) {
                cap = nf + 16;
                continue;
            }
            if (rcs[size_t(g)] == ZPAQGPU_OK) frs[size_t(g)].resize(size_t(nf));
            break;
        }
        m->stage_ms[size_t(g)] = float(now_ms() - t0);
    });
    if ((rc = first_error(m, rcs))) return rc;
    // 2. the exchange: one table over all devices, in file order
    struct Key {
        uint8_t b[24];
        bool operator==(const Key &o) const { return std::memcmp(b, o.b, 24) == 0; }
    };
    struct KeyHash {
        size_t operator()(const Key &k) const {
            uint64_t v;
            std::memcpy(&v, k.b, 8);
            return size_t(v);
        }
    };
    struct GFrag {
        u32 file, id, len;
        const uint8_t *sha1;
    };
    std::vector<GFrag> all;                       // every fragment, file order
    std::vector<std::vector<u32>> stored_of(static_cast<size_t>(G));  // per device: indices into `all` it stores
    std::unordered_map<Key, u32, KeyHash> seen;
    u32 next_id = 0;
    for (int g = 0; g < G; ++g)
        for (const zpaqgpu_fragment &f : frs[size_t(g)]) {
            GFrag e{f.file + u32(cut[size_t(g)]), 0, u32(f.len), f.sha1};
            bool first = true;
            if (opts->dedup) {
                Key k;
                std::memcpy(k.b, f.sha1, 20);
                const u32 l = u32(f.len);
                std::memcpy(k.b + 20, &l, 4);
                auto it = seen.find(k);
                if (it != seen.end()) e.id = it->second, first = false;
                else seen.emplace(k, next_id + 1);
            }
            if (first) {
                e.id = ++next_id;
                stored_of[size_t(g)].push_back(u32(all.size()));
            }
            all.push_back(e);
        }
    // 3. d blocks of every device: stored fragments in id order, packed up to block_bytes (jidac.v:186-214)
    struct DB {
        u32 first_id;
        u64 len;
        std::vector<u32> frags;  // indices into `all`
    };
    std::vector<std::vector<DB>> dbs(static_cast<size_t>(G));
    std::vector<u64> totals(size_t(G), 0);
    std::vector<std::vector<u64>> d_sizes(static_cast<size_t>(G));
    const std::vector<uint8_t> hdr = level_header(opts->level);
    Model model;
    if ((rc = model_from_level_layout(hdr.data(), int(hdr.size()), model))) return m->err = model.error, rc;
    // where each fragment's bytes are: fragment k of device g sits at in + frs[g][k].off (absolute in `in`? no:
    // zpaqgpu_jidac_fragment reports offsets inside `in`), so keep the source pointer beside the table
    std::vector<const uint8_t *> src_of(all.size());
    {
        size_t at = 0;
        for (int g = 0; g < G; ++g)
            for (const zpaqgpu_fragment &f : frs[size_t(g)]) src_of[at++] = in + f.off;
    }
    on_every_device(G, [&](int g) {
        std::vector<DB> &blocks = dbs[size_t(g)];
        for (u32 k : stored_of[size_t(g)]) {
            const GFrag &f = all[k];
            const bool fresh = blocks.empty() || opts->block_bytes == 0 || blocks.back().len + f.len > opts->block_bytes;
            if (fresh) blocks.push_back(DB{f.id, 0, {}});
            blocks.back().len += f.len, blocks.back().frags.push_back(k);
        }
        const int nd = int(blocks.size());
        if (nd == 0) return;
        const double t0 = now_ms();
        // the plaintext of the d blocks: the stored fragments back to back (duplicates skipped), gathered on
        // the host from the caller's buffer, then one upload
        u64 bytes = 0;
        for (const DB &b : blocks) bytes += b.len;
        std::vector<uint8_t> plain(static_cast<size_t>(bytes));
        std::vector<u64> off(size_t(nd) + 1, 0);
        std::vector<std::string> nm(static_cast<size_t>(nd)), cm(static_cast<size_t>(nd));
        std::vector<const char *> nmp(static_cast<size_t>(nd)), cmp(static_cast<size_t>(nd));
        u64 at = 0;
        for (int b = 0; b < nd; ++b) {
            off[size_t(b)] = at;
            for (u32 k : blocks[size_t(b)].frags) {
                if (all[k].len) std::memcpy(plain.data() + at, src_of[k], all[k].len);
                at += all[k].len;
            }
            nm[size_t(b)] = jidac_name(opts->date, 'd', blocks[size_t(b)].first_id);
            cm[size_t(b)] = jidac_comment(blocks[size_t(b)].len);
            nmp[size_t(b)] = nm[size_t(b)].c_str(), cmp[size_t(b)] = cm[size_t(b)].c_str();
        }
        off[size_t(nd)] = at;
        rcs[size_t(g)] = zg::guarded<int>(m->ctx[size_t(g)], [&]() -> int {
            int r = compress_stage(m->ctx[size_t(g)], model, plain.data(), off.data(), nd, nmp.data(), cmp.data(),
                                   &totals[size_t(g)]);
            if (r) return r;
            // the coded size of every d block goes into its h block
            d_sizes[size_t(g)].assign(size_t(nd) + 1, 0);
            zpaqgpu_ctx *c = m->ctx[size_t(g)];
            if (cudaSetDevice(c->device) != cudaSuccess) return ZPAQGPU_E_CUDA;
            if (cudaMemcpyAsync(d_sizes[size_t(g)].data(), c->out_off.p, 8 * size_t(nd + 1), cudaMemcpyDeviceToHost,
                                c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess)
                return ZPAQGPU_E_CUDA;
            return ZPAQGPU_OK;
        });
        m->stats[size_t(g)] = m->ctx[size_t(g)]->stats;
        m->stage_ms[size_t(g)] += float(now_ms() - t0);
    });
    if ((rc = first_error(m, rcs))) return rc;
    u64 total_d = 0;
    for (u64 t : totals) total_d += t;
    // 4. c, h and i blocks, all store mode (jidac.v:67-91, :216-295)
    std::vector<uint8_t> small;
    std::vector<u64> s_off{0};
    std::vector<std::string> s_name, s_comment;
    auto close_block = [&](char type, u32 num) {
        s_name.push_back(jidac_name(opts->date, type, num));
        s_comment.push_back(jidac_comment(small.size() - s_off.back()));
        s_off.push_back(small.size());
    };
    put_le(small, total_d, 8);
    close_block('c', next_id + 1);
    for (int g = 0; g < G; ++g)
        for (size_t b = 0; b < dbs[size_t(g)].size(); ++b) {
            const DB &d = dbs[size_t(g)][b];
            put_le(small, u32(d_sizes[size_t(g)][b + 1] - d_sizes[size_t(g)][b]), 4);
            for (u32 k : d.frags) {
                small.insert(small.end(), all[k].sha1, all[k].sha1 + 20);
                put_le(small, all[k].len, 4);
            }
            close_block('h', d.first_id);
        }
    {
        size_t fi = 0;
        for (int f = 0; f < n_files; ++f) {
            put_le(small, u64(opts->date), 8);
            for (const char *c = names[f] ? names[f] : ""; *c; ++c) small.push_back(uint8_t(*c));
            small.push_back(0);
            size_t fe = fi;
            while (fe < all.size() && all[fe].file == u32(f)) ++fe;
            if (opts->date != 0) {
                put_le(small, 0, 4);
                put_le(small, u32(fe - fi), 4);
                for (size_t k = fi; k < fe; ++k) put_le(small, all[k].id, 4);
            }
            fi = fe;
        }
        if (small.size() > s_off.back()) close_block('i', 1);
    }
    const int ns = int(s_name.size());
    std::vector<const char *> snp(static_cast<size_t>(ns)), scp(static_cast<size_t>(ns));
    for (int b = 0; b < ns; ++b) snp[size_t(b)] = s_name[size_t(b)].c_str(), scp[size_t(b)] = s_comment[size_t(b)].c_str();
    if (!m->aux && (rc = zpaqgpu_init(&m->aux, m->devices[0]))) return rc;
    const u64 cap2 = small.size() + small.size() / 4096 + 256 * u64(ns) + 4096;
    std::vector<uint8_t> idx(static_cast<size_t>(cap2));
    std::vector<u64> o2(size_t(ns) + 1, 0);
    u64 need2 = 0;
    rc = zpaqgpu_compress_blocks(m->aux, 0, small.data(), s_off.data(), ns, snp.data(), scp.data(), idx.data(), cap2,
                                 o2.data(), &need2);
    if (rc) return m->err = std::string("index blocks: ") + m->aux->err, rc;
    const u64 total_s = o2[size_t(ns)], c_len = o2[1], total = total_d + total_s;
    if (out_len) *out_len = total;
    if (out_need) *out_need = total;
    if (total > out_cap) return ZPAQGPU_E_NOSPACE;
    if (!out) return ZPAQGPU_E_ARG;
    // archive order: c block, d blocks (device order = id order), h blocks, i block (jidac.v:221-295)
    std::memcpy(out, idx.data(), size_t(c_len));
    std::vector<u64> base(size_t(G) + 1, c_len);
    for (int g = 0; g < G; ++g) base[size_t(g) + 1] = base[size_t(g)] + totals[size_t(g)];
    on_every_device(G, [&](int g) {
        const int nd = int(dbs[size_t(g)].size());
        if (nd == 0) return;
        const double t0 = now_ms();
        std::vector<u64> off(size_t(nd) + 1);
        rcs[size_t(g)] = zg::guarded<int>(m->ctx[size_t(g)], [&]() -> int {
            return compress_fetch(m->ctx[size_t(g)], nd, totals[size_t(g)], out + base[size_t(g)], off.data(), 0);
        });
        m->fetch_ms[size_t(g)] = float(now_ms() - t0);
    });
    if ((rc = first_error(m, rcs))) return rc;
    if (total_s > c_len) std::memcpy(out + c_len + total_d, idx.data() + c_len, size_t(total_s - c_len));
    return ZPAQGPU_OK;
    });
}

}  // extern "C"

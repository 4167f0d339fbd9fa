// zpaq_gpu.hpp -- C++ host mirror of the reference's zpaq.Compressor / zpaq.Decompresser / Reader /
// Writer (compressor.v:16-413, decompressor.v:170-640, io.v:6-21) over the C ABI of libzpaqgpu.
// Header-only; link with -lzpaqgpu.  Same method names, argument meaning and silent-error
// behaviour as the V types, so code written against them ports line for line.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zpaqgpu.h"

namespace zpaq {

struct Reader {                       // io.v:6-12
    virtual int get() = 0;            // one byte or -1 at EOF
    virtual ~Reader() = default;
};
struct Writer {                       // io.v:15-21
    virtual void put(int c) = 0;
    virtual void write(const uint8_t *p, size_t n) { for (size_t i = 0; i < n; ++i) put(p[i]); }
    virtual ~Writer() = default;
};
struct FileReader : Reader {          // io.v:41-78
    std::vector<uint8_t> data;
    size_t pos = 0;
    explicit FileReader(std::vector<uint8_t> d) : data(std::move(d)) {}
    int get() override { return pos < data.size() ? data[pos++] : -1; }
};
struct FileWriter : Writer {          // io.v:80-110
    std::vector<uint8_t> buf;
    void put(int c) override { buf.push_back(uint8_t(c)); }
    void write(const uint8_t *p, size_t n) override { buf.insert(buf.end(), p, p + n); }
    const std::vector<uint8_t> &bytes() const { return buf; }
};

class Gpu {                           // one context per GPU; no CPU fallback
  public:
    static zpaqgpu_ctx *ctx() {
        static Gpu g;
        return g.c_;
    }
  private:
    Gpu() {
        const int rc = zpaqgpu_init(&c_, -1);
        if (rc != ZPAQGPU_OK) throw std::runtime_error(zpaqgpu_strerror(rc));
    }
    ~Gpu() { zpaqgpu_destroy(c_); }
    zpaqgpu_ctx *c_ = nullptr;
};

// batch = true is for callers that write many blocks one after the other (cmd/main.v:288-317, one block per
// file): end_block() queues the block, the queue is coded in ONE launch when it is full or at flush() -- the one
// call such a caller adds before it closes its output.  Same bytes, same order on the Writer.
class Compressor {
  public:
    explicit Compressor(bool batch = false) : batch_(batch) {}
    void set_input(Reader *r) { in_ = r; }
    void set_output(Writer *w) { out_ = w; }
    void start_block(int level) {                                 // compressor.v:79
        if (state_ != Start) return;
        if (zpaqgpu_block_begin(Gpu::ctx(), level) == ZPAQGPU_OK) state_ = Block;
    }
    void start_segment(const std::string &filename, const std::string &comment) {   // compressor.v:212
        if (state_ != Block) return;
        if (zpaqgpu_segment_begin(Gpu::ctx(), filename.c_str(), comment.c_str()) == ZPAQGPU_OK) state_ = Segment;
    }
    bool compress(int n) {                                        // compressor.v:259
        if (state_ != Segment || !in_) return false;
        std::vector<uint8_t> buf;
        buf.reserve(size_t(n > 0 ? n : 0));
        while (int(buf.size()) < n) {
            const int c = in_->get();
            if (c < 0) break;
            buf.push_back(uint8_t(c));
        }
        zpaqgpu_segment_write(Gpu::ctx(), buf.data(), buf.size());
        return int(buf.size()) == n;
    }
    void end_segment() {                                          // compressor.v:357
        if (state_ != Segment) return;
        zpaqgpu_segment_end(Gpu::ctx());
        state_ = Block;
    }
    void end_block() {                                            // compressor.v:402
        if (state_ != Block) return;
        if (batch_) {
            const int full = zpaqgpu_block_end_queue(Gpu::ctx());
            state_ = Start;
            if (full == 1) flush();
            return;
        }
        uint64_t need = 0;
        zpaqgpu_block_end(Gpu::ctx(), nullptr, 0, &need);
        std::vector<uint8_t> blk(need);
        const int64_t n = zpaqgpu_block_end(Gpu::ctx(), blk.data(), need, &need);
        if (n > 0 && out_) out_->write(blk.data(), size_t(n));
        state_ = Start;
    }
    void flush() {                                                // not in the reference: delivers the queue
        if (state_ != Start) return;
        uint64_t need = 0;
        zpaqgpu_flush(Gpu::ctx(), nullptr, 0, &need);             // codes the queue; the bytes are kept
        if (!need) return;
        std::vector<uint8_t> all(need);
        const int64_t n = zpaqgpu_flush(Gpu::ctx(), all.data(), need, &need);
        if (n > 0 && out_) out_->write(all.data(), size_t(n));
    }
  private:
    enum { Block, Segment, Start } state_ = Start;                // compressor.v:6-8
    Reader *in_ = nullptr;
    Writer *out_ = nullptr;
    bool batch_ = false;
};

class Decompresser {
  public:
    void set_input(Reader *r) { in_ = r, loaded_ = false; }
    void set_output(Writer *w) { out_ = w; }
    bool find_block() {                                           // decompressor.v:219
        if (!in_) return false;
        load();
        while (cursor_ < segs_.size() && segs_[cursor_].block_index <= block_) ++cursor_;
        if (cursor_ >= segs_.size()) return false;
        block_ = segs_[cursor_].block_index;
        state_ = Block;
        return true;
    }
    bool find_filename() {                                        // decompressor.v:350
        if (state_ != Block) return false;
        if (cursor_ < segs_.size() && segs_[cursor_].block_index == block_) {
            cur_ = int(cursor_++), served_ = 0, state_ = Segment;
            return true;
        }
        state_ = Start;
        return false;
    }
    std::string get_filename() const { return cur_ < 0 ? "" : cstr(segs_[size_t(cur_)].name_off); }
    std::string get_comment() const { return cur_ < 0 ? "" : cstr(segs_[size_t(cur_)].comment_off); }
    bool decompress(int n = -1) {                                 // decompressor.v:443
        if (state_ != Segment) return false;
        const zpaqgpu_segment &s = segs_[size_t(cur_)];
        const uint64_t left = s.out_len - served_;
        const uint64_t take = (n < 0 || uint64_t(n) > left) ? left : uint64_t(n);
        if (take && out_) out_->write(plain_.data() + s.out_off + served_, size_t(take));
        served_ += take;
        return n >= 0 && take == uint64_t(n);
    }
    void read_segment_end() { if (state_ == Segment) state_ = Block; }   // decompressor.v:590
    int last_sha1_ok() const { return cur_ < 0 ? -1 : segs_[size_t(cur_)].sha1_ok; }
  private:
    void load() {
        if (loaded_) return;
        for (int c; (c = in_->get()) >= 0;) arc_.push_back(uint8_t(c));
        uint64_t cap = arc_.size() * 4 + 65536, need = 0;
        int seg_cap = 256, nseg = 0;
        for (;;) {
            plain_.resize(cap), segs_.resize(size_t(seg_cap));
            const int rc = zpaqgpu_decompress_archive(Gpu::ctx(), arc_.data(), arc_.size(), plain_.data(), cap, &need,
                                                      segs_.data(), seg_cap, &nseg);
            if (rc == ZPAQGPU_E_NOSPACE) { cap = need + 16, seg_cap = nseg + 16; continue; }
            break;
        }
        segs_.resize(size_t(nseg));
        loaded_ = true;
    }
    std::string cstr(uint64_t off) const {
        size_t e = size_t(off);
        while (e < arc_.size() && arc_[e]) ++e;
        return std::string(arc_.begin() + long(off), arc_.begin() + long(e));
    }
    enum { Block, Segment, Filename, Start } state_ = Start;      // decompressor.v:6-9
    Reader *in_ = nullptr;
    Writer *out_ = nullptr;
    std::vector<uint8_t> arc_, plain_;
    std::vector<zpaqgpu_segment> segs_;
    bool loaded_ = false;
    int block_ = -1, cur_ = -1;
    size_t cursor_ = 0;
    uint64_t served_ = 0;
};

// jidac.v:120-296 -- JidacArchive: create_archive writes c, d.., h.., i blocks to the Writer.
// `files` is an ordered list (a V map iterates in insertion order, which fixes the archive order).
class JidacArchive {
  public:
    explicit JidacArchive(int64_t date) : date_(date) {}          // date: get_jidac_date(), jidac.v:31-35
    void set_output(Writer *w) { out_ = w; }
    // the reference ignores `method` and always stores the d blocks (jidac.v:94-118)
    void create_archive(const std::vector<std::pair<std::string, std::vector<uint8_t>>> &files, int /*method*/) {
        run(files, 0, -1, false, 0);
    }
    // `jidac add`: rolling-hash fragments, SHA-1 dedup, stored fragments packed into coded d blocks
    void add(const std::vector<std::pair<std::string, std::vector<uint8_t>>> &files, int level, int fragment,
             uint64_t block_bytes) {
        run(files, level, fragment, true, block_bytes);
    }
  private:
    void run(const std::vector<std::pair<std::string, std::vector<uint8_t>>> &files, int level, int fragment,
             bool dedup, uint64_t block_bytes) {
        if (!out_) return;                                        // jidac.v:182-184
        std::vector<uint8_t> data;
        std::vector<uint64_t> off{0};
        std::vector<const char *> names;
        for (const auto &f : files) {
            data.insert(data.end(), f.second.begin(), f.second.end());
            off.push_back(data.size());
            names.push_back(f.first.c_str());
        }
        zpaqgpu_jidac_opts o{};
        o.date = date_, o.level = level, o.fragment = fragment, o.dedup = dedup ? 1 : 0, o.block_bytes = block_bytes;
        uint64_t got = 0, need = 0;
        std::vector<uint8_t> arc(data.size() + data.size() / 4 + 4096 * (files.size() + 4));
        int rc = zpaqgpu_jidac_add(Gpu::ctx(), &o, names.data(), data.data(), off.data(), int(files.size()), arc.data(),
                                   arc.size(), &got, &need);
        if (rc == ZPAQGPU_E_NOSPACE) {
            arc.resize(need);
            rc = zpaqgpu_jidac_add(Gpu::ctx(), &o, names.data(), data.data(), off.data(), int(files.size()), arc.data(),
                                   arc.size(), &got, &need);
        }
        if (rc == ZPAQGPU_OK) out_->write(arc.data(), size_t(got));
    }
    int64_t date_;
    Writer *out_ = nullptr;
};

}  // namespace zpaq

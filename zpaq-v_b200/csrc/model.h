// model.h -- host-side description of one ZPAQ model header: geometry, component table,
// per-block workspace layout in HBM and the constant lookup tables.
//
// Follows Compressor.start_block (compressor.v:96-145), Decompresser.find_block
// (decompressor.v:278-342) and Predictor.init (predictor.v:292-470) of the reference.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace zg {

// Component kinds, types.v:8-19.
enum : int { C_NONE = 0, C_CONS = 1, C_CM = 2, C_ICM = 3, C_MATCH = 4, C_AVG = 5, C_MIX2 = 6,
             C_MIX = 7, C_ISSE = 8, C_SSE = 9 };

// One component as the kernels see it.  Field meaning per type mirrors predictor.v:338-468.
struct CompDesc {
    int32_t type;
    int32_t a, b, c, limit;      // initial values of Component.a/.b/.c/.limit
    uint32_t p[4];               // MIX2: j,k,rate,mask   MIX: rate,mask
    uint64_t cm_off;             // byte offset of the u32 table inside the block workspace
    uint32_t cm_len;             // elements
    uint64_t ht_off;             // byte offset of the u8 table (hash slots / MATCH buffer)
    uint32_t ht_len;             // bytes
    uint64_t a16_off;            // MIX2 weights (u16)
    uint32_t a16_len;
    int32_t level;               // 0: needs no other component's prediction; else 1 + the deepest input in
                                 // front of it (warp kernel: components of one level are evaluated together)
};

constexpr uint32_t kPageBytes = 256;   // 4 hash lines of 64 bytes

struct FillRegion {              // workspace words that do not start as zero
    uint64_t off;                // byte offset inside the block workspace (4-aligned)
    uint64_t n_words;
    uint32_t img_off;            // first word of the repeating image in Model::image
    uint32_t period;             // image length in words
};

enum : int { CTX_VM = 0, CTX_M1 = 1, CTX_HASHCHAIN = 2 };

struct Model {
    std::vector<uint8_t> header; // z.header
    int cend = 0, hbegin = 0, hend = 0;
    int n = 0;                   // number of components (header[4])
    std::vector<CompDesc> comps;
    // ZPAQL VM memory (zpaql.v:74-96): H words, M bytes, R[256]
    uint64_t h_off = 0, m_off = 0, r_off = 0;
    uint64_t rt_off = 0;         // per-component runtime state of the generic kernel (n x 32 B)
    uint32_t h_len = 0, m_len = 0;
    uint64_t ws_bytes = 0;       // workspace stride per block
    std::vector<FillRegion> fills;
    // Paged layout (chain-shaped models only): the ICM/ISSE hash tables are virtual; per block only a
    // page table (one u32 per 256-byte page) is resident and pages come from a pool shared by the wave.
    std::vector<CompDesc> comps_paged;   // ht_off = offset of the page table, ht_len = virtual length
    std::vector<FillRegion> fills_paged;
    uint64_t ws_bytes_paged = 0;
    std::vector<uint32_t> image;
    // bytes start_block writes: locator "zPQ" lvl 1 hsize COMP HCOMP (compressor.v:150-181)
    std::vector<uint8_t> block_prefix;
    // specialised-kernel shape: ICM, n_isse x ISSE(j=i-1), optional MIX2(j=n-3,k=n-2)
    bool is_chain = false;
    int n_isse = 0;
    bool has_mix2 = false;
    int ctx_mode = CTX_VM;
    int n_hash = 0;              // CTX_HASHCHAIN: number of HASH/*d=a rounds
    std::string error;
};

// The lookup tables built once per process (predictor.v:21-106, statetable.v:15-100).
struct Tables {
    int32_t squash[4096];
    int32_t stretch[32768];
    int32_t dt[1024];
    int32_t dt2k[256];
    uint8_t ns[1024];
};
const Tables &tables();
int st_cminit(int state);
int h_stretch(int p);
int h_squash(int d);
int h_clamp512k(int x);

// get_compression_level(level).hcomp, levels.v:26-375
std::vector<uint8_t> level_header(int level);

// Geometry from the levels.v layout, the way start_block derives it.  Returns a ZPAQGPU_* code.
int model_from_level_layout(const uint8_t *hdr, int len, Model &m);
// Geometry from archive bytes the way find_block reads them: `p` points at the level byte just
// after the locator.  *consumed = bytes up to and including the HCOMP terminator.
int model_from_archive(const uint8_t *p, uint64_t avail, Model &m, uint64_t *consumed);

} // namespace zg

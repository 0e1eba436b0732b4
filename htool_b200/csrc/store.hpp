// htool_b200/csrc/store.hpp — the device-resident flattened leaf store: formats shared by the host
// packer (store.cpp) and the sm_100a kernels (kernels.cu).
//
// Reference data model being re-laid-out (read-only, via htb_leaf): dense leaves are m x n column-major
// (include/htool/matrix/matrix.hpp:20-26), low-rank leaves are U (m x r column-major) and V (r x n
// column-major) (include/htool/hmatrix/lrmat/lrmat.hpp:16-54).
//
// Layout in HBM ("stream-ordered slabs"):
//   * two index spaces ("sides"): side 0 = target rows, side 1 = source columns. Each side is cut into
//     BLOCKS of <= block_rows consecutive indices, cut points chosen on leaf boundaries;
//   * side 0 holds the U panels (m x r) and the dense leaves (m x n); side 1 holds the transposed V panels
//     Vt = V^T (n x r), so every panel is "long dimension fast". A panel is cut by the side's blocks into
//     CHUNKS (h x W, h <= block_rows), a chunk into UNITS of <= unit_elems coefficients (h x w, w <= 32),
//     stored column-major with leading dimension h;
//   * all units of one block are concatenated, in leaf order, into the block's STREAM, itself cut into
//     STAGES (<= stage_bytes, 16 B aligned) = one bulk-async copy each. A stage starts with its own
//     unit descriptors, so a CTA needs nothing but the stream: [StageHeader | Unit x n | coefficients].
//   Every coefficient is stored exactly once per side it is needed on and is read exactly once per
//   pass; rank and size bucketing happen through the (h, w) unit shape, not through separate arrays.
//
// Two generic passes run over a side (kernels.cu):
//   REDUCE  partial[k] = sum_i op(P[i,k]) * in[i]     (t = V x, or t = op(U)^T x / z = op(A)^T x)
//   APPLY   out[i]    += sum_k op(P[i,k]) * c[k]      (U t, A x, or op(Vt) t; + ADDVEC z)
// and every product of the reference (N / T / C, plain / symmetric-twice) is a fixed sequence of them.
#ifndef HTB_STORE_HPP
#define HTB_STORE_HPP

#include <cstddef>
#include <cstdint>
#include <vector>

#ifdef __CUDACC__
#    define HTB_HD __host__ __device__
#else
#    define HTB_HD
#endif

namespace htb {

enum UnitKind : uint32_t { UNIT_LOWRANK = 0, // panel of U (side 0) or of Vt (side 1)
                           UNIT_DENSE   = 1, // panel of a dense leaf (side 0 only)
                           UNIT_ADDVEC  = 2  // side 1 only, no coefficients: out[i] += z[i] (dense leaf, transposed application)
};

// geom: row0 [0:8) | h-1 [8:16) | w [16:24) | kind [24:26) | applied_twice [26]
struct Unit {
    uint32_t data_off;   // element offset of the h x w panel inside the stage's coefficient region
    uint32_t geom;
    uint32_t aux_apply;  // LOWRANK: scratch offset of t[k0..k0+w). DENSE: first input index (leaf col_offset + k0). ADDVEC: scratch offset of z for row0
    uint32_t aux_reduce; // LOWRANK/DENSE: scratch offset receiving this unit's w partial sums
};
static_assert(sizeof(Unit) == 16, "Unit must be 16 bytes");

HTB_HD inline uint32_t unit_row0(uint32_t g) { return g & 0xffu; }
HTB_HD inline uint32_t unit_h(uint32_t g) { return ((g >> 8) & 0xffu) + 1u; }
HTB_HD inline uint32_t unit_w(uint32_t g) { return (g >> 16) & 0xffu; }
HTB_HD inline uint32_t unit_kind(uint32_t g) { return (g >> 24) & 0x3u; }
HTB_HD inline uint32_t unit_twice(uint32_t g) { return (g >> 26) & 0x1u; }
inline uint32_t make_geom(uint32_t row0, uint32_t h, uint32_t w, uint32_t kind, uint32_t twice) {
    return (row0 & 0xffu) | (((h - 1u) & 0xffu) << 8) | ((w & 0xffu) << 16) | ((kind & 3u) << 24) | ((twice & 1u) << 26);
}

struct StageHeader {
    uint32_t n_units;
    uint32_t data_byte_off; // from the start of the stage to the coefficient region (multiple of 16)
    uint32_t reserved[2];
};
static_assert(sizeof(StageHeader) == 16, "StageHeader must be 16 bytes");

struct StageDesc {
    uint64_t byte_off; // into the side's stream, multiple of 16
    uint32_t nbytes;   // multiple of 16
    uint32_t flags;    // bit 0: holds at least one applied-twice unit
};

struct BlockDesc {
    int32_t row_start; // first index of the block in the side's index space
    int32_t nrows;
    uint32_t first_stage;
    uint32_t n_stages;
    uint32_t flags; // bit 0: holds at least one applied-twice unit
    uint32_t reserved[3];
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc must be 32 bytes");

// t = sum of the per-chunk partial vectors of a leaf that spans several blocks on the reduce side
struct CombineEntry {
    uint32_t dst;      // scratch offset of the final vector
    uint32_t src;      // scratch offset of the first partial (partials are consecutive, w apart)
    uint32_t w;        // vector length (rank, or nb_cols for a dense leaf's z)
    uint32_t n_chunks; // bit 31: leaf is applied twice
};

struct PackOptions {
    int block_rows  = 64;    // <= 128, multiple of 32
    int unit_elems  = 512;   // coefficients per unit
    int stage_bytes = 16384; // bulk-copy granule, multiple of 16
};

// Host description of one side, produced by the packer. Device copies are owned by the handle.
struct SideLayout {
    int n = 0; // length of the index space
    std::vector<BlockDesc> blocks;
    std::vector<StageDesc> stages;
    std::vector<uint32_t> order; // block ids, heaviest stream first
    std::vector<CombineEntry> combine;
    uint64_t stream_bytes = 0;
    uint64_t n_units      = 0;
    bool any_twice        = false;
};

} // namespace htb
#endif

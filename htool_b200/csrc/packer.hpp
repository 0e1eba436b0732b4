// htool_b200/csrc/packer.hpp — host-side packer: htb_leaf list -> stream-ordered store (see store.hpp).
// Runs once per assembly (htb_create). Plain C++17 + OpenMP, no CUDA: it only produces bytes and tables;
// capi.cu uploads them.
#ifndef HTB_PACKER_HPP
#define HTB_PACKER_HPP

#include "store.hpp"
#include <algorithm>
#include <htool_b200.h>
#include <string>
#include <vector>

namespace htb {

class Packer {
  public:
    // Plans the layout (blocks, units, stages, c-stream slots, partial areas, combine tables).
    // Throws std::runtime_error on invalid input.
    Packer(const htb_hmatrix_desc &desc, const PackOptions &opt);

    // Writes the stream bytes of blocks [b0, b1) of side s to dst, dst[0] being the first byte of block
    // b0's stream (blocks are contiguous in the side's stream, in block order).
    void fill(int s, int b0, int b1, char *dst) const;
    // Device assembly: no leaf carries host coefficients (all_on_device), so a stream is zeros except for the headers of its
    // stages (StageHeader + Unit table, 16 + 16 n_units bytes each). fill_headers writes those of blocks [b0, b1) back to
    // back: the header of stage st starts at header_offset(s, st) - header_offset(s, first stage of b0).
    bool all_on_device = false;
    void fill_headers(int s, int b0, int b1, char *dst) const;
    uint64_t header_offset(int s, size_t st) const { return m_hdr_off[s][st]; } // st in [0, n_stages]
    const std::vector<uint64_t> &header_offsets(int s) const { return m_hdr_off[s]; }

    uint64_t block_offset(int s, int b) const { return m_block_off[s][b]; } // byte offset of block b's stream, b in [0, nblocks]

    int dtype;
    size_t esize;
    PackOptions opt;
    int piece = 16; // effective piece_cols
    SideLayout side[2];
    NearFieldLayout nf; // multi-RHS near field of side 0 (store.hpp)
    uint64_t scratch_elems = 0; // elements of one scratch copy: PART[0] | PART[1] | CS[0] | CS[1]
    uint64_t tf_elems = 0, mscratch_elems = 0; // multi-RHS scratch copy, in vectors: TF | PARTM[0] | PARTM[1]

    // statistics
    int64_t n_leaves = 0, n_dense = 0, n_lowrank = 0, n_twice = 0, coefficients = 0, coefficients_twice = 0;
    int rank_min = 0, rank_max = 0;
    int nb_rows = 0, nb_cols = 0;

  private:
    static constexpr uint32_t kDirect = 0xffffffffu;
    struct UnitSpec;
    const htb_leaf *m_leaves;
    std::vector<uint64_t> m_piece_ptr;     // per leaf: first global piece index (n_leaves + 1)
    std::vector<int32_t> m_block_start[2]; // nblocks + 1
    std::vector<int32_t> m_blk_of[2];      // index -> block
    std::vector<int32_t> m_first_blk[2];   // per leaf: first block it touches on side s
    std::vector<int32_t> m_nchunks[2];     // per leaf: number of blocks it touches on side s (0: leaf holds nothing)
    std::vector<uint64_t> m_chunk_ptr[2];  // per leaf: prefix sum of m_nchunks (n_leaves + 1)
    RawVector<uint64_t> m_inc_index[2];    // per (leaf, chunk): position of the incidence in m_csr_leaf
    std::vector<uint64_t> m_csr_ptr[2];    // nblocks + 1
    RawVector<uint32_t> m_csr_leaf[2];     // incidences (leaf ids), block-major, leaf order inside a block
    std::vector<uint64_t> m_unit_ptr[2];   // per incidence: first unit index (n_incidences + 1)
    RawVector<uint32_t> m_unit_slot[2];    // per unit: scratch offset of its c slot
    RawVector<uint16_t> m_unit_cslot[2];   // per unit: offset of its c slot inside its stage's c segment
    std::vector<uint32_t> m_part_off[2];   // [consumer side] per global piece: scratch offset of the partials, or kDirect
    std::vector<uint64_t> m_block_off[2];
    std::vector<uint64_t> m_hdr_off[2];    // prefix sums of the stages' header bytes (n_stages + 1)

    void prefetch_leaf(int s, uint32_t li) const { // what a walk over a block's incidences reads per leaf
        __builtin_prefetch(&m_leaves[li]);
        __builtin_prefetch(&m_first_blk[s][li]);
        __builtin_prefetch(&m_piece_ptr[li]);
    }
    bool active(const htb_leaf &l) const { return l.nb_rows > 0 && l.nb_cols > 0 && l.rank != 0; }
    int vec_len(const htb_leaf &l) const { return l.rank < 0 ? l.nb_cols : l.rank; }
    int n_pieces(const htb_leaf &l) const { return (vec_len(l) + piece - 1) / piece; }
    int piece_len(const htb_leaf &l, int p) const { return std::min(piece, vec_len(l) - p * piece); }
    int start_of(int s, const htb_leaf &l) const { return s == 0 ? l.row_offset : l.col_offset; }
    int extent_of(int s, const htb_leaf &l) const { return s == 0 ? l.nb_rows : l.nb_cols; }
    // [lo, hi) of chunk c of leaf li along side s, relative to the leaf
    void chunk_range(int s, uint32_t li, int c, int &lo, int &hi) const;
    // ADDVEC units (side 1, dense leaf): pieces [p_lo, p_hi] met by chunk c
    void addvec_pieces(uint32_t li, int c, int &p_lo, int &p_hi) const;
    int units_in_incidence(int s, uint32_t li, int c) const;
    uint64_t unit_of(int s, uint32_t li, int c, int j) const { return m_unit_ptr[s][m_inc_index[s][m_chunk_ptr[s][li] + c]] + j; }
    uint32_t producer_out(int ps, const UnitSpec &u) const;

    void make_blocks(int s);
    void make_incidence_lists(int s);
    void make_incidence(int s);
    void make_partials(int cs);
    void make_tf();
    void make_partm_sizes();
    void make_partm(int cs);
    uint64_t m_partm_total[2] = {0, 0};
    void make_combine(int cs);
    void make_mtables();
    std::vector<uint32_t> m_partm_off[2]; // [consumer side] per global piece: PARTM offset (vectors) or kDirect
    std::vector<uint32_t> m_tf_off;       // per global piece: TF offset (vectors)
    template <typename Emit>
    void walk_block(int s, int b, Emit &&emit) const;
    void layout_block(int s, int b, std::vector<StageDesc> &stages, RawVector<uint32_t> &unit_stage, uint64_t &n_units, bool &any_twice);
    template <typename T>
    void fill_block(int s, int b, char *dst, bool headers_only = false) const;
    struct NfSrc { // a dense unit of side 0 in the main stream
        uint64_t src_off;
        uint32_t col; // first column (root numbering)
        uint16_t row0, h, w, ld;
    };
    void make_near_field(const std::vector<std::vector<NfSrc>> &per_block);
};

} // namespace htb
#endif

// htool_b200/csrc/packer.hpp — host-side packer: htb_leaf list -> stream-ordered store (see store.hpp).
// Runs once per assembly (htb_create). Plain C++17 + OpenMP, no CUDA: it only produces bytes; capi.cu
// uploads them.
#ifndef HTB_PACKER_HPP
#define HTB_PACKER_HPP

#include "store.hpp"
#include <htool_b200.h>
#include <string>
#include <vector>

namespace htb {

class Packer {
  public:
    // Plans the layout (blocks, units, stages, scratch offsets). Throws std::runtime_error on invalid input.
    Packer(const htb_hmatrix_desc &desc, const PackOptions &opt);

    // Writes the stream bytes of blocks [b0, b1) of side s to dst, dst[0] being the first byte of block
    // b0's stream (blocks are contiguous in the side's stream, in block order).
    void fill(int s, int b0, int b1, char *dst) const;

    uint64_t block_offset(int s, int b) const { return m_block_off[s][b]; } // byte offset of block b's stream, b in [0, nblocks]

    int dtype;
    size_t esize;
    PackOptions opt;
    SideLayout side[2];
    uint64_t scratch_elems = 0; // elements of one scratch copy: final t/z vectors + per-chunk partials of both sides

    // statistics
    int64_t n_leaves = 0, n_dense = 0, n_lowrank = 0, n_twice = 0, coefficients = 0, coefficients_twice = 0;
    int rank_min = 0, rank_max = 0;
    int nb_rows = 0, nb_cols = 0;

  private:
    struct Ref { // one (leaf, block) incidence
        uint32_t leaf;
    };
    const htb_leaf *m_leaves;
    std::vector<uint32_t> m_toff;        // per leaf: scratch offset of its final vector
    std::vector<uint32_t> m_pbase[2];    // per leaf: scratch offset of its partials on side s (== m_toff when single chunk)
    std::vector<int32_t> m_first_blk[2]; // per leaf: first block it touches on side s
    std::vector<int32_t> m_nchunks[2];
    std::vector<int32_t> m_block_start[2]; // nblocks + 1
    std::vector<uint64_t> m_csr_ptr[2];    // nblocks + 1
    std::vector<uint32_t> m_csr_leaf[2];
    std::vector<uint64_t> m_block_off[2];

    void make_blocks(int s);
    void make_incidence(int s);
    template <typename Emit>
    void walk_block(int s, int b, Emit &&emit) const;
    void layout_block(int s, int b, std::vector<StageDesc> &stages, uint64_t &n_units, bool &any_twice) const;
    template <typename T>
    void fill_block(int s, int b, char *dst) const;
};

} // namespace htb
#endif

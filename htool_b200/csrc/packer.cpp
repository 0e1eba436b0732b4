// htool_b200/csrc/packer.cpp — see packer.hpp / store.hpp.
#include "packer.hpp"

#include <algorithm>
#include <complex>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace htb {

namespace {

struct UnitSpec {
    uint32_t leaf;
    uint32_t row0, h; // rows [row0, row0+h) of the block
    uint32_t p0;      // first panel row of the chunk (offset inside the leaf along this side)
    uint32_t k0, w;   // panel columns [k0, k0+w)
    uint32_t kind, twice;
    uint32_t aux_apply, aux_reduce;
    uint32_t elems() const { return kind == UNIT_ADDVEC ? 0u : h * w; }
};

inline uint32_t round16(uint64_t v) { return static_cast<uint32_t>((v + 15u) & ~uint64_t(15)); }

// Greedy stage cutter shared by the layout pass and the fill pass so both see the same stages.
struct StageCutter {
    size_t esize;
    uint32_t stage_bytes;
    uint32_t nu = 0;
    uint64_t data_elems = 0;
    bool fits(uint32_t elems) const {
        if (nu == 0)
            return true;
        return 16u + 16u * (nu + 1u) + (data_elems + elems) * esize <= stage_bytes;
    }
    void add(uint32_t elems) {
        nu++;
        data_elems += elems;
    }
    uint32_t header_bytes() const { return 16u + 16u * nu; }
    uint32_t nbytes() const { return round16(header_bytes() + data_elems * esize); }
    void reset() {
        nu         = 0;
        data_elems = 0;
    }
};

} // namespace

Packer::Packer(const htb_hmatrix_desc &desc, const PackOptions &o) : dtype(desc.dtype), esize(desc.dtype == HTB_DOUBLE ? 8 : 16), opt(o), m_leaves(desc.leaves) {
    if (desc.dtype != HTB_DOUBLE && desc.dtype != HTB_COMPLEX_DOUBLE)
        throw std::runtime_error("unknown dtype");
    if (desc.nb_rows < 0 || desc.nb_cols < 0 || desc.nb_leaves < 0 || (desc.nb_leaves > 0 && !desc.leaves))
        throw std::runtime_error("invalid H-matrix description");
    if (opt.block_rows < 32 || opt.block_rows > 128 || opt.block_rows % 32)
        throw std::runtime_error("block_rows must be 32, 64, 96 or 128");
    if (opt.unit_elems < 32 || opt.stage_bytes % 16 || static_cast<size_t>(opt.stage_bytes) < 48 + static_cast<size_t>(std::max(opt.unit_elems, 32 * opt.block_rows / 32)) * 16)
        throw std::runtime_error("invalid unit_elems / stage_bytes");
    nb_rows  = desc.nb_rows;
    nb_cols  = desc.nb_cols;
    n_leaves = desc.nb_leaves;

    // validation + statistics
    rank_min = INT32_MAX;
    rank_max = -1;
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (l.nb_rows < 0 || l.nb_cols < 0 || l.row_offset < 0 || l.col_offset < 0 || l.row_offset + l.nb_rows > nb_rows || l.col_offset + l.nb_cols > nb_cols)
            throw std::runtime_error("leaf " + std::to_string(i) + " lies outside the root block");
        bool empty = l.nb_rows == 0 || l.nb_cols == 0;
        if (l.rank < 0) {
            n_dense++;
            if (!empty && !l.data0)
                throw std::runtime_error("dense leaf without data");
            if ((l.flags & (HTB_LEAF_DIAG_SYMMETRIC | HTB_LEAF_DIAG_HERMITIAN)) && l.nb_rows != l.nb_cols)
                throw std::runtime_error("symmetric dense leaf is not square");
            coefficients += int64_t(l.nb_rows) * l.nb_cols;
        } else {
            n_lowrank++;
            if (!empty && l.rank > 0 && (!l.data0 || !l.data1))
                throw std::runtime_error("low-rank leaf without data");
            rank_min = std::min(rank_min, l.rank);
            rank_max = std::max(rank_max, l.rank);
            coefficients += int64_t(l.rank) * (l.nb_rows + l.nb_cols);
        }
        if (l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) {
            n_twice++;
            coefficients_twice += l.rank < 0 ? int64_t(l.nb_rows) * l.nb_cols : int64_t(l.rank) * (l.nb_rows + l.nb_cols);
        }
    }
    if (n_lowrank == 0)
        rank_min = 0;

    side[0].n = nb_rows;
    side[1].n = nb_cols;
    for (int s = 0; s < 2; s++)
        make_blocks(s);
    for (int s = 0; s < 2; s++)
        make_incidence(s);

    // scratch layout: [final vectors | side-0 partials | side-1 partials]
    m_toff.assign(n_leaves, 0);
    uint64_t off = 0;
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        m_toff[i]         = static_cast<uint32_t>(off);
        off += l.rank < 0 ? l.nb_cols : l.rank;
    }
    for (int s = 0; s < 2; s++) {
        m_pbase[s].assign(n_leaves, 0);
        for (int64_t i = 0; i < n_leaves; i++) {
            const htb_leaf &l = m_leaves[i];
            uint64_t W        = l.rank < 0 ? (s == 0 ? l.nb_cols : 0) : l.rank; // dense leaves only reduce on side 0
            if (m_nchunks[s][i] > 1 && W > 0) {
                m_pbase[s][i] = static_cast<uint32_t>(off);
                side[s].combine.push_back(CombineEntry{m_toff[i], static_cast<uint32_t>(off), static_cast<uint32_t>(W), static_cast<uint32_t>(m_nchunks[s][i]) | ((l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) ? 0x80000000u : 0u)});
                off += W * static_cast<uint64_t>(m_nchunks[s][i]);
            } else {
                m_pbase[s][i] = m_toff[i];
            }
            if (off >= (uint64_t(1) << 32))
                throw std::runtime_error("scratch exceeds 2^32 elements");
        }
    }
    scratch_elems = off;

    // per-block stage layout
    for (int s = 0; s < 2; s++) {
        const int nb = static_cast<int>(side[s].blocks.size());
        std::vector<std::vector<StageDesc>> per_block(nb);
        std::vector<uint64_t> units(nb, 0);
        std::vector<char> twice(nb, 0);
#pragma omp parallel for schedule(dynamic, 16)
        for (int b = 0; b < nb; b++) {
            bool t = false;
            layout_block(s, b, per_block[b], units[b], t);
            twice[b] = t;
        }
        m_block_off[s].assign(nb + 1, 0);
        uint64_t stream = 0;
        size_t nstages  = 0;
        for (int b = 0; b < nb; b++)
            nstages += per_block[b].size();
        side[s].stages.reserve(nstages);
        for (int b = 0; b < nb; b++) {
            m_block_off[s][b]       = stream;
            BlockDesc &bd           = side[s].blocks[b];
            bd.first_stage          = static_cast<uint32_t>(side[s].stages.size());
            bd.n_stages             = static_cast<uint32_t>(per_block[b].size());
            bd.flags                = twice[b] ? 1u : 0u;
            side[s].any_twice       = side[s].any_twice || twice[b];
            side[s].n_units += units[b];
            for (StageDesc sd : per_block[b]) {
                sd.byte_off += stream;
                side[s].stages.push_back(sd);
            }
            if (!per_block[b].empty())
                stream = side[s].stages.back().byte_off + side[s].stages.back().nbytes;
        }
        m_block_off[s][nb]  = stream;
        side[s].stream_bytes = stream;
        // heaviest blocks first: the hardware block scheduler then balances the tail
        side[s].order.resize(nb);
        std::iota(side[s].order.begin(), side[s].order.end(), 0u);
        std::stable_sort(side[s].order.begin(), side[s].order.end(), [&](uint32_t a, uint32_t b) {
            return (m_block_off[s][a + 1] - m_block_off[s][a]) > (m_block_off[s][b + 1] - m_block_off[s][b]);
        });
    }
}

// Cut [0, n) into blocks of <= block_rows indices. Cut points are taken where no small leaf (<= block_rows
// long on this side) is split, whenever such a point exists within reach: for the nested ranges of a
// cluster tree this yields cluster-aligned blocks, so small leaves are never split.
void Packer::make_blocks(int s) {
    const int n  = side[s].n;
    const int BR = opt.block_rows;
    std::vector<int32_t> cost(static_cast<size_t>(n) + 2, 0);
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (l.nb_rows == 0 || l.nb_cols == 0 || l.rank == 0)
            continue;
        int a = s == 0 ? l.row_offset : l.col_offset, len = s == 0 ? l.nb_rows : l.nb_cols;
        if (len >= 2 && len <= BR) {
            cost[a + 1]++;
            cost[a + len]--;
        }
    }
    for (int i = 1; i <= n; i++)
        cost[i] += cost[i - 1];
    std::vector<int32_t> &start = m_block_start[s];
    start.clear();
    start.push_back(0);
    int c = 0;
    while (c < n) {
        int limit = std::min(n, c + BR);
        int p     = limit;
        if (limit < n) {
            for (int q = limit; q > c; q--)
                if (cost[q] == 0) {
                    p = q;
                    break;
                }
        }
        start.push_back(p);
        c = p;
    }
    const int nb = static_cast<int>(start.size()) - 1;
    side[s].blocks.assign(nb, BlockDesc{});
    for (int b = 0; b < nb; b++) {
        side[s].blocks[b].row_start = start[b];
        side[s].blocks[b].nrows     = start[b + 1] - start[b];
    }
}

void Packer::make_incidence(int s) {
    const int n  = side[s].n;
    const int nb = static_cast<int>(side[s].blocks.size());
    std::vector<int32_t> blk_of(static_cast<size_t>(n) + 1, 0);
    for (int b = 0; b < nb; b++)
        for (int i = m_block_start[s][b]; i < m_block_start[s][b + 1]; i++)
            blk_of[i] = b;
    m_first_blk[s].assign(n_leaves, 0);
    m_nchunks[s].assign(n_leaves, 0);
    m_csr_ptr[s].assign(static_cast<size_t>(nb) + 1, 0);
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (l.nb_rows == 0 || l.nb_cols == 0 || l.rank == 0)
            continue;
        int a = s == 0 ? l.row_offset : l.col_offset, len = s == 0 ? l.nb_rows : l.nb_cols;
        int b0 = blk_of[a], b1 = blk_of[a + len - 1];
        m_first_blk[s][i] = b0;
        m_nchunks[s][i]   = b1 - b0 + 1;
        for (int b = b0; b <= b1; b++)
            m_csr_ptr[s][b + 1]++;
    }
    for (int b = 0; b < nb; b++)
        m_csr_ptr[s][b + 1] += m_csr_ptr[s][b];
    m_csr_leaf[s].assign(m_csr_ptr[s][nb], 0);
    std::vector<uint64_t> cursor(m_csr_ptr[s].begin(), m_csr_ptr[s].end() - 1);
    for (int64_t i = 0; i < n_leaves; i++) // leaf order is kept inside every block: fixed summation order
        for (int c = 0; c < m_nchunks[s][i]; c++)
            m_csr_leaf[s][cursor[m_first_blk[s][i] + c]++] = static_cast<uint32_t>(i);
}

template <typename Emit>
void Packer::walk_block(int s, int b, Emit &&emit) const {
    const int bs = m_block_start[s][b], be = m_block_start[s][b + 1];
    for (uint64_t e = m_csr_ptr[s][b]; e < m_csr_ptr[s][b + 1]; e++) {
        const uint32_t li = m_csr_leaf[s][e];
        const htb_leaf &l = m_leaves[li];
        const int a = s == 0 ? l.row_offset : l.col_offset, len = s == 0 ? l.nb_rows : l.nb_cols;
        const int lo = std::max(a, bs), hi = std::min(a + len, be);
        UnitSpec u{};
        u.leaf  = li;
        u.row0  = static_cast<uint32_t>(lo - bs);
        u.h     = static_cast<uint32_t>(hi - lo);
        u.p0    = static_cast<uint32_t>(lo - a);
        u.twice = (l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) ? 1u : 0u;
        const uint32_t chunk = static_cast<uint32_t>(b - m_first_blk[s][li]);
        if (l.rank < 0 && s == 1) {
            u.kind       = UNIT_ADDVEC;
            u.k0         = 0;
            u.w          = 0;
            u.aux_apply  = m_toff[li] + u.p0;
            u.aux_reduce = 0;
            emit(u);
            continue;
        }
        const uint32_t W    = l.rank < 0 ? static_cast<uint32_t>(l.nb_cols) : static_cast<uint32_t>(l.rank);
        const uint32_t wmax = std::min<uint32_t>(32u, std::max<uint32_t>(1u, static_cast<uint32_t>(opt.unit_elems) / u.h));
        u.kind              = l.rank < 0 ? UNIT_DENSE : UNIT_LOWRANK;
        const uint32_t red0 = m_pbase[s][li] + (m_nchunks[s][li] > 1 ? chunk * W : 0u);
        for (uint32_t k0 = 0; k0 < W; k0 += wmax) {
            u.k0         = k0;
            u.w          = std::min(wmax, W - k0);
            u.aux_apply  = l.rank < 0 ? static_cast<uint32_t>(l.col_offset) + k0 : m_toff[li] + k0;
            u.aux_reduce = red0 + k0;
            emit(u);
        }
    }
}

void Packer::layout_block(int s, int b, std::vector<StageDesc> &stages, uint64_t &n_units, bool &any_twice) const {
    StageCutter cut{esize, static_cast<uint32_t>(opt.stage_bytes)};
    uint64_t off     = 0;
    uint32_t stflags = 0;
    auto close       = [&]() {
        if (cut.nu == 0)
            return;
        stages.push_back(StageDesc{off, cut.nbytes(), stflags});
        off += cut.nbytes();
        cut.reset();
        stflags = 0;
    };
    walk_block(s, b, [&](const UnitSpec &u) {
        if (!cut.fits(u.elems()))
            close();
        cut.add(u.elems());
        n_units++;
        if (u.twice) {
            stflags |= 1u;
            any_twice = true;
        }
    });
    close();
}

namespace {
template <typename T>
inline T conj_of(T v);
template <>
inline double conj_of(double v) { return v; }
template <>
inline std::complex<double> conj_of(std::complex<double> v) { return std::conj(v); }
template <typename T>
inline T real_of(T v);
template <>
inline double real_of(double v) { return v; }
template <>
inline std::complex<double> real_of(std::complex<double> v) { return std::complex<double>(v.real(), 0.); }
} // namespace

template <typename T>
void Packer::fill_block(int s, int b, char *dst) const {
    StageCutter cut{esize, static_cast<uint32_t>(opt.stage_bytes)};
    std::vector<UnitSpec> pending;
    char *cursor = dst;
    auto close   = [&]() {
        if (cut.nu == 0)
            return;
        const uint32_t nbytes = cut.nbytes();
        StageHeader hdr{cut.nu, cut.header_bytes(), {0, 0}};
        std::memcpy(cursor, &hdr, sizeof(hdr));
        Unit *units = reinterpret_cast<Unit *>(cursor + sizeof(StageHeader));
        T *data     = reinterpret_cast<T *>(cursor + cut.header_bytes());
        uint32_t eoff = 0;
        for (uint32_t i = 0; i < cut.nu; i++) {
            const UnitSpec &u = pending[i];
            units[i]          = Unit{eoff, make_geom(u.row0, u.h, u.w, u.kind, u.twice), u.aux_apply, u.aux_reduce};
            if (u.kind == UNIT_ADDVEC)
                continue;
            const htb_leaf &l = m_leaves[u.leaf];
            T *out            = data + eoff;
            if (u.kind == UNIT_LOWRANK && s == 1) {
                // Vt panel: element (i, k) = V[k + (p0+i) * r], V is r x n column-major
                const T *V  = static_cast<const T *>(l.data1);
                const size_t r = static_cast<size_t>(l.rank);
                for (uint32_t k = 0; k < u.w; k++)
                    for (uint32_t i = 0; i < u.h; i++)
                        out[i + static_cast<size_t>(k) * u.h] = V[(u.k0 + k) + (u.p0 + i) * r];
            } else if (u.kind == UNIT_DENSE && (l.flags & (HTB_LEAF_DIAG_SYMMETRIC | HTB_LEAF_DIAG_HERMITIAN))) {
                // symv / hemv read only the UPLO triangle (add_matrix_vector_product.hpp:26-52): rebuild the full
                // block from that triangle so the kernels see an ordinary dense leaf
                const T *A     = static_cast<const T *>(l.data0);
                const size_t m = static_cast<size_t>(l.nb_rows);
                const bool upper = (l.flags & HTB_LEAF_UPLO_UPPER) != 0;
                const bool herm  = (l.flags & HTB_LEAF_DIAG_HERMITIAN) != 0;
                for (uint32_t k = 0; k < u.w; k++)
                    for (uint32_t i = 0; i < u.h; i++) {
                        const size_t gi = u.p0 + i, gj = u.k0 + k;
                        const bool stored = upper ? gi <= gj : gi >= gj;
                        T v;
                        if (gi == gj)
                            v = herm ? real_of<T>(A[gi + gi * m]) : A[gi + gi * m];
                        else if (stored)
                            v = A[gi + gj * m];
                        else
                            v = herm ? conj_of<T>(A[gj + gi * m]) : A[gj + gi * m];
                        out[i + static_cast<size_t>(k) * u.h] = v;
                    }
            } else {
                // U panel or dense leaf: column-major with lda = nb_rows
                const T *A     = static_cast<const T *>(l.data0);
                const size_t m = static_cast<size_t>(l.nb_rows);
                for (uint32_t k = 0; k < u.w; k++)
                    std::memcpy(out + static_cast<size_t>(k) * u.h, A + u.p0 + (u.k0 + k) * m, sizeof(T) * u.h);
            }
            eoff += u.elems();
        }
        const size_t used = cut.header_bytes() + static_cast<size_t>(cut.data_elems) * esize;
        if (used < nbytes)
            std::memset(cursor + used, 0, nbytes - used);
        cursor += nbytes;
        cut.reset();
        pending.clear();
    };
    walk_block(s, b, [&](const UnitSpec &u) {
        if (!cut.fits(u.elems()))
            close();
        cut.add(u.elems());
        pending.push_back(u);
    });
    close();
}

void Packer::fill(int s, int b0, int b1, char *dst) const {
    const uint64_t base = m_block_off[s][b0];
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = b0; b < b1; b++) {
        char *p = dst + (m_block_off[s][b] - base);
        if (dtype == HTB_DOUBLE)
            fill_block<double>(s, b, p);
        else
            fill_block<std::complex<double>>(s, b, p);
    }
}

} // namespace htb

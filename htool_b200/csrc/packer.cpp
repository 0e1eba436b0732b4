// htool_b200/csrc/packer.cpp — see packer.hpp / store.hpp.
#include "packer.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <cstring>
#include <exception>
#include <numeric>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace htb {

struct Packer::UnitSpec {
    uint32_t leaf;
    uint64_t ui;      // unit index on its side
    int chunk;        // chunk index of the leaf on this side
    int piece;        // piece index of the leaf
    uint32_t row0, h; // rows [row0, row0+h) of the block
    uint32_t p0;      // offset of the unit's first row inside the leaf (along this side)
    uint32_t k0, w;   // panel columns [k0, k0+w) (0, 0 for ADDVEC)
    uint32_t sub_off; // ADDVEC: offset of the unit inside its piece
    uint32_t kind, twice;
    uint32_t ld; // leading dimension of the stored panel
    uint32_t elems() const { return kind == UNIT_ADDVEC ? 0u : ld * w; }
    uint32_t celems() const { return kind == UNIT_ADDVEC ? h : w; }
};

namespace {

inline uint32_t round16(uint64_t v) { return static_cast<uint32_t>((v + 15u) & ~uint64_t(15)); }

// Greedy stage cutter shared by the layout pass and the fill pass so both see the same stages.
struct StageCutter {
    size_t esize;
    uint32_t stage_bytes, cseg_bytes;
    uint32_t nu         = 0;
    uint64_t data_elems = 0, c_elems = 0;
    bool fits(uint32_t elems, uint32_t celems) const {
        if (nu == 0)
            return true;
        return 16u + 16u * (nu + 1u) + (data_elems + elems) * esize <= stage_bytes && (c_elems + celems) * esize <= cseg_bytes;
    }
    void add(uint32_t elems, uint32_t celems) {
        nu++;
        data_elems += elems;
        c_elems += celems;
    }
    uint32_t header_bytes() const { return 16u + 16u * nu; }
    uint32_t nbytes() const { return round16(header_bytes() + data_elems * esize); }
    uint32_t c_len_padded() const { return static_cast<uint32_t>(esize == 8 ? (c_elems + 1u) & ~uint64_t(1) : c_elems); }
    void reset() {
        nu         = 0;
        data_elems = 0;
        c_elems    = 0;
    }
};

// The serial passes over the leaf list are independent of each other (per side, per direction, per table): job j of n runs on
// thread j of a team of the usual size (a team of n would make libgomp drop and re-create its other threads around every
// call). f must not open a parallel region of its own (it would be serialised); exceptions leave through the calling thread.
template <typename F>
void concurrently(int n, F &&f) {
    std::vector<std::exception_ptr> err(n);
    auto run = [&](int j) {
        try {
            f(j);
        } catch (...) {
            err[j] = std::current_exception();
        }
    };
#ifdef _OPENMP
#pragma omp parallel
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        for (int j = t; j < n; j += nt)
            run(j);
    }
#else
    for (int j = 0; j < n; j++)
        run(j);
#endif
    for (const std::exception_ptr &e : err)
        if (e)
            std::rethrow_exception(e);
}
template <typename F>
void both_sides(F &&f) {
    concurrently(2, f);
}

} // namespace

Packer::Packer(const htb_hmatrix_desc &desc, const PackOptions &o) : dtype(desc.dtype), esize(desc.dtype == HTB_DOUBLE ? 8 : 16), opt(o), m_leaves(desc.leaves) {
    if (desc.dtype != HTB_DOUBLE && desc.dtype != HTB_COMPLEX_DOUBLE)
        throw std::runtime_error("unknown dtype");
    if (desc.nb_rows < 0 || desc.nb_cols < 0 || desc.nb_leaves < 0 || (desc.nb_leaves > 0 && !desc.leaves))
        throw std::runtime_error("invalid H-matrix description");
    if (opt.block_rows == 0)
        opt.block_rows = esize == 8 ? 128 : 64;
    if ((opt.block_rows != 32 && opt.block_rows != 64 && opt.block_rows != 128) || static_cast<size_t>(opt.block_rows) * esize > 1024)
        throw std::runtime_error("block_rows must be 32, 64 or 128 (32 or 64 for complex)");
    if (opt.target_block_rows != 0 && opt.target_block_rows != 32 && opt.target_block_rows != 64 && opt.target_block_rows != 128)
        throw std::runtime_error("target_block_rows must be 0 (automatic), 32, 64 or 128");
    if (opt.stage_bytes % 16 || opt.cseg_bytes % 16 || opt.stage_bytes > 65536 || opt.piece_cols < 1 || opt.piece_cols > 32)
        throw std::runtime_error("invalid piece_cols / stage_bytes / cseg_bytes");
    // a unit (block_rows x piece) and its descriptor must fit one stage; its c vector one c segment
    piece = 1;
    const size_t max_ld = unit_ld(static_cast<uint32_t>(opt.block_rows), esize); // (the tallest unit has the largest leading dimension)
    while (piece * 2 <= opt.piece_cols && 32u + max_ld * (piece * 2) * esize <= static_cast<size_t>(opt.stage_bytes))
        piece *= 2;
    if (32u + max_ld * piece * esize > static_cast<size_t>(opt.stage_bytes))
        throw std::runtime_error("stage_bytes too small for one unit");
    if (static_cast<size_t>(std::max(opt.block_rows, piece)) * esize > static_cast<size_t>(opt.cseg_bytes) || opt.cseg_bytes / esize > 65535)
        throw std::runtime_error("cseg_bytes too small for one unit (or too large)");
    nb_rows  = desc.nb_rows;
    nb_cols  = desc.nb_cols;
    n_leaves = desc.nb_leaves;
    if (n_leaves >= (int64_t(1) << 32))
        throw std::runtime_error("too many leaves");

    // validation + statistics
    rank_min = INT32_MAX;
    rank_max = -1;
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (l.nb_rows < 0 || l.nb_cols < 0 || l.row_offset < 0 || l.col_offset < 0 || l.row_offset + l.nb_rows > nb_rows || l.col_offset + l.nb_cols > nb_cols)
            throw std::runtime_error("leaf " + std::to_string(i) + " lies outside the root block");
        bool empty = l.nb_rows == 0 || l.nb_cols == 0;
        if (l.rank < -1)
            throw std::runtime_error("leaf " + std::to_string(i) + " is an admissible block still to be compressed (rank HTB_RANK_COMPRESS): use htb_create_compressed");
        if (l.rank < 0) {
            n_dense++;
            if (!empty && !l.data0 && !opt.generate_dense)
                throw std::runtime_error("dense leaf without data");
            if ((l.flags & (HTB_LEAF_DIAG_SYMMETRIC | HTB_LEAF_DIAG_HERMITIAN)) && l.nb_rows != l.nb_cols)
                throw std::runtime_error("symmetric dense leaf is not square");
            coefficients += int64_t(l.nb_rows) * l.nb_cols;
        } else {
            n_lowrank++;
            if (!empty && l.rank > 0 && (!l.data0 || !l.data1) && !(opt.generate_dense && !l.data0 && !l.data1))
                throw std::runtime_error("low-rank leaf without data"); // (both pointers null + a generator: factors in the device pool, aca.cu)
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

    // HTB_PACK_TIMING=1: seconds of every phase of the layout on stderr (development aid)
    const bool timing = std::getenv("HTB_PACK_TIMING") != nullptr;
    auto t_last       = std::chrono::steady_clock::now();
    auto lap          = [&](const char *what) {
        if (!timing)
            return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[htb pack] %-14s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    lap("validation");
    m_piece_ptr.assign(n_leaves + 1, 0);
    for (int64_t i = 0; i < n_leaves; i++)
        m_piece_ptr[i + 1] = m_piece_ptr[i] + (active(m_leaves[i]) ? n_pieces(m_leaves[i]) : 0);

    side[0].n = nb_rows;
    side[1].n = nb_cols;
    both_sides([&](int s) { make_blocks(s); });
    lap("make_blocks");
    for (int s = 0; s < 2; s++)
        make_incidence_lists(s);
    lap("incidence lists");
    for (int s = 0; s < 2; s++)
        make_incidence(s);
    lap("make_incidence");
    // three serial passes over the leaf list at once: the partial areas of the two directions and the TF offsets of the multi-RHS path
    concurrently(3, [&](int j) {
        if (j < 2)
            make_partials(j);
        else
            make_tf();
    });
    lap("make_partials");

    // per-block stage layout
    RawVector<uint32_t> unit_stage[2]; // (every entry is written by the walks: no serial zero fill of tens of MB)
    for (int s = 0; s < 2; s++) {
        const int nb = static_cast<int>(side[s].blocks.size());
        const uint64_t nu = m_unit_ptr[s].back();
        unit_stage[s].resize(nu);
        m_unit_cslot[s].resize(nu);
        m_unit_slot[s].resize(nu);
        std::vector<std::vector<StageDesc>> per_block(nb);
        std::vector<uint64_t> units(nb, 0);
        std::vector<char> twice(nb, 0);
#pragma omp parallel for schedule(dynamic, 16)
        for (int b = 0; b < nb; b++) {
            bool t = false;
            layout_block(s, b, per_block[b], unit_stage[s], units[b], t);
            twice[b] = t;
        }
        lap(s == 0 ? "  layout walk 0" : "  layout walk 1");
        // where every block's stages, stream bytes, c segments and aux records start (prefix sums over the blocks), then the
        // side-wide stage table in parallel
        std::vector<uint64_t> stage_at(static_cast<size_t>(nb) + 1, 0), c_at(static_cast<size_t>(nb) + 1, 0), aux_at(static_cast<size_t>(nb) + 1, 0);
        m_block_off[s].assign(nb + 1, 0);
        for (int b = 0; b < nb; b++) {
            uint64_t bytes = 0, c_len = 0, aux = 0;
            for (const StageDesc &sd : per_block[b]) {
                bytes += sd.nbytes;
                c_len += sd.c_len;
                aux += sd.aux_off16; // (layout_block left the size of the stage's aux record here)
            }
            stage_at[b + 1]       = stage_at[b] + per_block[b].size();
            m_block_off[s][b + 1] = m_block_off[s][b] + bytes;
            c_at[b + 1]           = c_at[b] + c_len;
            aux_at[b + 1]         = aux_at[b] + aux;
            side[s].any_twice     = side[s].any_twice || twice[b];
            side[s].n_units += units[b];
        }
        const uint64_t stream = m_block_off[s][nb], c_off = c_at[nb], aux_off = aux_at[nb];
        if (aux_off / 16u >= (uint64_t(1) << 32))
            throw std::runtime_error("multi-RHS aux tables too large");
        side[s].stages.resize(stage_at[nb]);
        uint32_t aux_max  = 0;
        bool aux_overflow = false;
#pragma omp parallel for schedule(static) reduction(max : aux_max) reduction(|| : aux_overflow)
        for (int b = 0; b < nb; b++) {
            BlockDesc &bd     = side[s].blocks[b];
            bd.first_stage    = static_cast<uint32_t>(stage_at[b]);
            bd.n_stages       = static_cast<uint32_t>(per_block[b].size());
            bd.flags          = twice[b] ? 1u : 0u;
            bd.n_twice_stages = 0;
            uint64_t c = c_at[b], aux = aux_at[b];
            size_t at  = stage_at[b];
            for (StageDesc sd : per_block[b]) {
                bd.n_twice_stages += sd.flags & 1u;
                sd.byte_off += m_block_off[s][b];
                sd.first_unit += static_cast<uint32_t>(m_unit_ptr[s][m_csr_ptr[s][b]]); // block-local -> side-wide
                sd.c_off = static_cast<uint32_t>(c);
                c += sd.c_len;
                const uint32_t aux_bytes = sd.aux_off16;
                // (cannot happen: a stage holds <= cseg_bytes / esize columns)
                aux_overflow = aux_overflow || aux_bytes > aux_slot_bytes(static_cast<uint32_t>(opt.cseg_bytes)) || aux_bytes / 16u > 0x7fffu;
                sd.aux_off16 = static_cast<uint32_t>(aux / 16u);
                sd.flags     = static_cast<uint16_t>((sd.flags & 1u) | ((aux_bytes / 16u) << 1));
                aux += aux_bytes;
                aux_max               = std::max(aux_max, aux_bytes);
                side[s].stages[at++] = sd;
            }
            std::vector<StageDesc>().swap(per_block[b]);
        }
        if (aux_overflow)
            throw std::runtime_error("multi-RHS aux record exceeds its slot");
        side[s].aux_max_bytes = std::max(side[s].aux_max_bytes, aux_max);
        lap(s == 0 ? "  layout stages 0" : "  layout stages 1");
        side[s].aux_reduce.resize(aux_off); // written by make_mtables
        side[s].aux_apply.resize(aux_off);
        side[s].stream_bytes = stream;
        side[s].cs_elems     = c_off;
        // heaviest blocks first: the hardware block scheduler then balances the tail
        side[s].order.resize(nb);
        std::iota(side[s].order.begin(), side[s].order.end(), 0u);
        std::stable_sort(side[s].order.begin(), side[s].order.end(), [&](uint32_t a, uint32_t b) {
            return (m_block_off[s][a + 1] - m_block_off[s][a]) > (m_block_off[s][b + 1] - m_block_off[s][b]);
        });
    }

    lap("layout_blocks");
    // scratch copy: [PART[0] | PART[1] | CS[0] | CS[1]], c-stream bases 16 B aligned
    side[0].part_base = 0;
    side[1].part_base = side[0].part_elems;
    side[0].cs_base   = (side[1].part_base + side[1].part_elems + 1u) & ~uint64_t(1);
    side[1].cs_base   = side[0].cs_base + side[0].cs_elems;
    scratch_elems     = (side[1].cs_base + side[1].cs_elems + 1u) & ~uint64_t(1); // even: a second copy stays 16 B aligned
    if (scratch_elems >= (uint64_t(1) << 32))
        throw std::runtime_error("scratch exceeds 2^32 elements");
    for (int s = 0; s < 2; s++) {
        const int64_t n_part = static_cast<int64_t>(m_part_off[s].size());
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_part; i++)
            if (m_part_off[s][i] != kDirect)
                m_part_off[s][i] += static_cast<uint32_t>(side[s].part_base);
        const int nb = static_cast<int>(side[s].blocks.size());
#pragma omp parallel for schedule(dynamic, 64)
        for (int b = 0; b < nb; b++) {
            const uint64_t u0 = m_unit_ptr[s][m_csr_ptr[s][b]], u1 = m_unit_ptr[s][m_csr_ptr[s][b + 1]];
            for (uint64_t ui = u0; ui < u1; ui++)
                m_unit_slot[s][ui] = static_cast<uint32_t>(side[s].cs_base + side[s].stages[side[s].blocks[b].first_stage + unit_stage[s][ui]].c_off + m_unit_cslot[s][ui]);
        }
    }
    bool some_host_data = false;
#pragma omp parallel for schedule(static) reduction(|| : some_host_data)
    for (int64_t i = 0; i < n_leaves; i++)
        some_host_data = some_host_data || (active(m_leaves[i]) && m_leaves[i].data0);
    all_on_device = n_leaves > 0 && !some_host_data;
    for (int s = 0; s < 2; s++) {
        m_hdr_off[s].assign(side[s].stages.size() + 1, 0);
        for (size_t st = 0; st < side[s].stages.size(); st++)
            m_hdr_off[s][st + 1] = m_hdr_off[s][st] + 16u + 16u * side[s].stages[st].n_units;
    }
    lap("unit_slots");
    // four more: the combine tables of the two directions, the multi-RHS partial areas of the two directions
    make_partm_sizes();
    concurrently(4, [&](int j) {
        if (j < 2)
            make_combine(j);
        else
            make_partm(j - 2);
    });
    lap("make_combine");
    make_mtables();
    lap("make_mtables");
}

// Cut [0, n) into blocks of <= block_rows indices. Cut points are taken where no small leaf (<= block_rows
// long on this side) is split, whenever such a point exists within reach: for the nested ranges of a
// cluster tree this yields cluster-aligned blocks, so small leaves are never split.
void Packer::make_blocks(int s) {
    const int n = side[s].n;
    int BR      = opt.block_rows;
    if (s == 0 && opt.target_block_rows > 0)
        BR = std::min(opt.block_rows, opt.target_block_rows);
    std::vector<int32_t> &start = m_block_start[s];
    // appends the cut points of [from, to) for blocks of <= br indices
    auto cut_range = [&](int from, int to, int br) {
        std::vector<int32_t> cost(static_cast<size_t>(n) + 2, 0);
        for (int64_t i = 0; i < n_leaves; i++) {
            const htb_leaf &l = m_leaves[i];
            if (!active(l))
                continue;
            int a = start_of(s, l), len = extent_of(s, l);
            if (len >= 2 && len <= br) {
                cost[a + 1]++;
                cost[a + len]--;
            }
        }
        for (int i = 1; i <= n; i++)
            cost[i] += cost[i - 1];
        int c = from;
        while (c < to) {
            int limit = std::min(to, c + br);
            int p     = limit;
            if (limit < to) {
                for (int q = limit; q > c; q--)
                    if (cost[q] == 0) {
                        p = q;
                        break;
                    }
            }
            start.push_back(p);
            c = p;
        }
    };
    start.clear();
    start.push_back(0);
    cut_range(0, n, BR);
    // Tail split (side 0 of a small shard, e.g. a row strip of a distributed operator): APPLY blocks are nearly uniform,
    // so nb blocks over `cta_slots` resident CTAs run as floor(nb / slots) full rounds plus a last round of r blocks
    // during which most SMs idle (1024 blocks on 444 slots: 2.31 rounds, the last one at a third of the bandwidth).
    // The rows of those last r blocks are re-cut into quarter-height blocks: they sort last in the heaviest-first launch
    // order and fill every slot, so the tail streams at full rate.
    const int slots = opt.cta_slots, small = std::max(32, BR / 4);
    int nb = static_cast<int>(start.size()) - 1;
    if (s == 0 && opt.tail_split && slots > 0 && nb > slots && nb <= 6 * slots && small < BR) {
        const int r = nb % slots;
        if (r > 0 && r * 5 <= slots * 4) {
            const int tail_from = start[nb - r];
            start.resize(static_cast<size_t>(nb - r) + 1);
            cut_range(tail_from, n, small);
            nb = static_cast<int>(start.size()) - 1;
        }
    }
    side[s].blocks.assign(nb, BlockDesc{});
    m_blk_of[s].assign(static_cast<size_t>(n) + 1, 0);
    for (int b = 0; b < nb; b++) {
        side[s].blocks[b].row_start = start[b];
        side[s].blocks[b].nrows     = start[b + 1] - start[b];
        for (int i = start[b]; i < start[b + 1]; i++)
            m_blk_of[s][i] = b;
    }
}

void Packer::chunk_range(int s, uint32_t li, int c, int &lo, int &hi) const {
    const htb_leaf &l = m_leaves[li];
    const int a = start_of(s, l), len = extent_of(s, l), b = m_first_blk[s][li] + c;
    lo = std::max(a, m_block_start[s][b]) - a;
    hi = std::min(a + len, m_block_start[s][b + 1]) - a;
}

void Packer::addvec_pieces(uint32_t li, int c, int &p_lo, int &p_hi) const {
    int lo, hi;
    chunk_range(1, li, c, lo, hi);
    p_lo = lo / piece;
    p_hi = (hi - 1) / piece;
}

int Packer::units_in_incidence(int s, uint32_t li, int c) const {
    const htb_leaf &l = m_leaves[li];
    if (l.rank < 0 && s == 1) {
        int p_lo, p_hi;
        addvec_pieces(li, c, p_lo, p_hi);
        return p_hi - p_lo + 1;
    }
    return n_pieces(l);
}

// The incidence structure of side s: which leaf meets which block, block-major, leaf order kept inside a block.
void Packer::make_incidence_lists(int s) {
    const int nb = static_cast<int>(side[s].blocks.size());
    m_first_blk[s].assign(n_leaves, 0);
    m_nchunks[s].assign(n_leaves, 0);
    m_chunk_ptr[s].assign(n_leaves + 1, 0);
    m_csr_ptr[s].assign(static_cast<size_t>(nb) + 1, 0);
    // A counting sort of the (leaf, block) incidences by block, in parallel over contiguous ranges of the leaf list: thread t
    // counts its range per block, the counts are scanned block-major then thread-major, and every thread scatters its range
    // from its own cursors — inside a block the leaves of thread t follow those of thread t - 1, i.e. leaf order is kept
    // exactly as a serial pass would keep it (fixed summation order, whatever the number of threads).
#ifdef _OPENMP
    const int max_threads = std::max(1, omp_get_max_threads());
#else
    const int max_threads = 1;
#endif
    std::vector<std::vector<uint64_t>> cursor(max_threads);
    std::vector<uint64_t> chunks_before(static_cast<size_t>(max_threads) + 1, 0);
#pragma omp parallel num_threads(max_threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int t = 0, nt = 1;
#endif
        const int64_t lo = n_leaves * t / nt, hi = n_leaves * (t + 1) / nt;
        std::vector<uint64_t> &cnt = cursor[t];
        cnt.assign(nb, 0);
        uint64_t my_chunks = 0;
        for (int64_t i = lo; i < hi; i++) {
            const htb_leaf &l = m_leaves[i];
            if (!active(l))
                continue;
            const int a = start_of(s, l), len = extent_of(s, l);
            const int b0 = m_blk_of[s][a], b1 = m_blk_of[s][a + len - 1];
            m_first_blk[s][i] = b0;
            m_nchunks[s][i]   = b1 - b0 + 1;
            my_chunks += static_cast<uint64_t>(b1 - b0 + 1);
            for (int b = b0; b <= b1; b++)
                cnt[b]++;
        }
        chunks_before[t + 1] = my_chunks;
#pragma omp barrier
#pragma omp single
        {
            for (int u = 0; u < nt; u++)
                chunks_before[u + 1] += chunks_before[u];
            m_chunk_ptr[s][n_leaves] = chunks_before[nt];
            uint64_t at = 0;
            for (int b = 0; b < nb; b++) {
                m_csr_ptr[s][b] = at;
                for (int u = 0; u < nt; u++) { // count -> first position of thread u's leaves inside block b
                    const uint64_t c = cursor[u][b];
                    cursor[u][b]     = at;
                    at += c;
                }
            }
            m_csr_ptr[s][nb] = at;
            m_csr_leaf[s].resize(at); // (default-initialised: every incidence is written exactly once below)
            m_inc_index[s].resize(at);
        } // (implicit barrier)
        uint64_t run = chunks_before[t];
        for (int64_t i = lo; i < hi; i++) {
            m_chunk_ptr[s][i] = run;
            for (int c = 0; c < m_nchunks[s][i]; c++) {
                const uint64_t e         = cnt[m_first_blk[s][i] + c]++;
                m_csr_leaf[s][e]         = static_cast<uint32_t>(i);
                m_inc_index[s][run + c] = e;
            }
            run += static_cast<uint64_t>(m_nchunks[s][i]);
        }
    }
}

// Parallel part: order of the incidences inside every block, unit counts.
void Packer::make_incidence(int s) {
    const int nb         = static_cast<int>(side[s].blocks.size());
    const uint64_t n_inc = m_csr_ptr[s][nb];
    const bool tm = std::getenv("HTB_PACK_TIMING") != nullptr;
    auto t0       = std::chrono::steady_clock::now();
    auto lapi     = [&](const char *w) { if (tm) { auto n = std::chrono::steady_clock::now(); std::fprintf(stderr, "[htb pack]   incidence(%d) %-10s %.3f s\n", s, w, std::chrono::duration<double>(n - t0).count()); t0 = n; } };
    // Inside a block, order the incidences by (first row, height, applied-twice): the panels acting on the same rows of
    // the block become consecutive in the stream and form RUNS (store.hpp) for the multi-RHS kernels. The sort is stable,
    // so leaf order — hence the summation order — stays fixed inside a run.
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < (opt.sort_units ? nb : 0); b++) {
        const uint64_t e0 = m_csr_ptr[s][b], e1 = m_csr_ptr[s][b + 1];
        auto key = [&](uint32_t li) {
            int lo, hi;
            chunk_range(s, li, b - m_first_blk[s][li], lo, hi);
            const htb_leaf &l = m_leaves[li];
            const uint64_t r0 = static_cast<uint64_t>(start_of(s, l) + lo - m_block_start[s][b]);
            // (low-rank panels before dense leaves inside a group: the dense columns are the tail of a run, RunDesc::K_lr)
            // (only when the near-field layout is wanted: leaf order inside a group is otherwise kept as it is — measured, APPLY_M is
            // 2 % faster with the dense columns interleaved with the low-rank ones than with all of them at the end of a run)
            return (r0 << 32) | (static_cast<uint64_t>(hi - lo) << 2) | ((l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) ? 2u : 0u) | ((opt.near_field && l.rank < 0) ? 1u : 0u);
        };
        // (keys computed once per incidence: a key reads the leaf record, a cache miss in a list of millions of leaves)
        std::vector<std::pair<uint64_t, uint32_t>> keyed(e1 - e0);
        for (uint64_t e = e0; e < e1; e++) {
            if (e + 8 < e1)
                prefetch_leaf(s, m_csr_leaf[s][e + 8]);
            keyed[e - e0] = {key(m_csr_leaf[s][e]), m_csr_leaf[s][e]};
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<uint64_t, uint32_t> &a, const std::pair<uint64_t, uint32_t> &c) { return a.first < c.first; });
        for (uint64_t e = e0; e < e1; e++) {
            const uint32_t li                                              = keyed[e - e0].second;
            m_csr_leaf[s][e]                                               = li;
            m_inc_index[s][m_chunk_ptr[s][li] + (b - m_first_blk[s][li])] = e;
        }
    }
    lapi("sort");
    m_unit_ptr[s].assign(n_inc + 1, 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < nb; b++)
        for (uint64_t e = m_csr_ptr[s][b]; e < m_csr_ptr[s][b + 1]; e++) {
            if (e + 8 < m_csr_ptr[s][b + 1])
                prefetch_leaf(s, m_csr_leaf[s][e + 8]);
            const uint32_t li    = m_csr_leaf[s][e];
            m_unit_ptr[s][e + 1] = static_cast<uint64_t>(units_in_incidence(s, li, b - m_first_blk[s][li]));
        }
    for (uint64_t e = 0; e < n_inc; e++)
        m_unit_ptr[s][e + 1] += m_unit_ptr[s][e];
    lapi("units");
}

// For every piece and direction (= consumer side cs): can its single producer write straight into its single
// consumer slot, or does it go through partials + COMBINE? Offsets are relative to PART[cs] here.
void Packer::make_partials(int cs) {
    {
        const int ps = 1 - cs;
        m_part_off[cs].assign(m_piece_ptr[n_leaves], kDirect);
        uint64_t off = 0;
        for (int64_t i = 0; i < n_leaves; i++) {
            const htb_leaf &l = m_leaves[i];
            if (!active(l))
                continue;
            const uint32_t li = static_cast<uint32_t>(i);
            for (int p = 0; p < n_pieces(l); p++) {
                const int len = piece_len(l, p);
                int n_prod = m_nchunks[ps][i], n_cons = m_nchunks[cs][i];
                bool sum_kind = true;
                if (l.rank < 0) {
                    // the ADDVEC units of side 1 that meet this piece
                    const int a  = l.col_offset;
                    const int c0 = m_blk_of[1][a + p * piece] - m_first_blk[1][li];
                    const int c1 = m_blk_of[1][a + p * piece + len - 1] - m_first_blk[1][li];
                    if (cs == 0) {
                        n_prod   = c1 - c0 + 1;
                        sum_kind = false; // x slices are copied, not summed
                    } else {
                        n_cons = c1 - c0 + 1;
                    }
                }
                if (n_prod == 1 && n_cons == 1)
                    continue;
                m_part_off[cs][m_piece_ptr[i] + p] = static_cast<uint32_t>(off);
                off += static_cast<uint64_t>(sum_kind ? n_prod : 1) * len;
                if (off >= (uint64_t(1) << 32))
                    throw std::runtime_error("scratch exceeds 2^32 elements");
            }
        }
        side[cs].part_elems = off;
    }
}

void Packer::make_combine(int cs) {
    const int ps = 1 - cs;
    SideLayout &sl = side[cs];
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (!active(l))
            continue;
        const uint32_t li = static_cast<uint32_t>(i);
        for (int p = 0; p < n_pieces(l); p++) {
            const uint32_t po = m_part_off[cs][m_piece_ptr[i] + p];
            if (po == kDirect)
                continue;
            const int len        = piece_len(l, p);
            const bool sum_kind  = !(l.rank < 0 && cs == 0);
            const uint32_t n_sum = sum_kind ? static_cast<uint32_t>(m_nchunks[ps][i]) : 1u;
            if (n_sum >= (1u << 24))
                throw std::runtime_error("too many chunks in one leaf");
            CombineEntry ce{po, static_cast<uint32_t>(sl.combine_dst.size()), 0u, n_sum | (static_cast<uint32_t>(len) << 24) | ((l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) ? 0x80000000u : 0u)};
            if (l.rank < 0 && cs == 1) {
                // consumers are the ADDVEC units that meet the piece, each wants its sub-range of z
                const int a  = l.col_offset;
                const int c0 = m_blk_of[1][a + p * piece] - m_first_blk[1][li];
                const int c1 = m_blk_of[1][a + p * piece + len - 1] - m_first_blk[1][li];
                for (int c = c0; c <= c1; c++) {
                    int lo, hi, p_lo, p_hi;
                    chunk_range(1, li, c, lo, hi);
                    addvec_pieces(li, c, p_lo, p_hi);
                    const int sub_lo = std::max(lo, p * piece), sub_hi = std::min(hi, p * piece + len);
                    sl.combine_dst.push_back(CombineDst{m_unit_slot[1][unit_of(1, li, c, p - p_lo)], static_cast<uint16_t>(sub_lo - p * piece), static_cast<uint16_t>(sub_hi - sub_lo)});
                }
            } else {
                for (int c = 0; c < m_nchunks[cs][i]; c++)
                    sl.combine_dst.push_back(CombineDst{m_unit_slot[cs][unit_of(cs, li, c, p)], 0, static_cast<uint16_t>(len)});
            }
            ce.n_dst = static_cast<uint32_t>(sl.combine_dst.size()) - ce.dst_first;
            sl.combine.push_back(ce);
        }
    }
}

// Side tables of the multi-RHS path (store.hpp, MUnit): TF offsets per piece, partial areas for pieces with several
// producer chunks, one MUnit per unit in stage order, and the panel-buffer batches of every stage.
void Packer::make_tf() {
    const uint64_t n_pieces_total = m_piece_ptr[n_leaves];
    m_tf_off.assign(n_pieces_total, 0);
    uint64_t tf = 0;
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (!active(l))
            continue;
        for (int p = 0; p < n_pieces(l); p++) {
            m_tf_off[m_piece_ptr[i] + p] = static_cast<uint32_t>(tf);
            tf += piece_len(l, p);
        }
    }
    tf_elems = tf;
}

// Sizes of the two multi-RHS partial areas (a parallel sum), so that both directions can be laid out concurrently (make_partm).
void Packer::make_partm_sizes() {
    for (int cs = 0; cs < 2; cs++) {
        uint64_t sum = 0;
#pragma omp parallel for schedule(static) reduction(+ : sum)
        for (int64_t i = 0; i < n_leaves; i++) {
            const htb_leaf &l = m_leaves[i];
            if (active(l) && !(l.rank < 0 && cs == 0) && m_nchunks[1 - cs][i] > 1)
                sum += static_cast<uint64_t>(m_nchunks[1 - cs][i]) * static_cast<uint64_t>(vec_len(l));
        }
        m_partm_total[cs] = sum;
    }
}

// Partial area of direction cs (consumer side cs) of the multi-RHS path, after make_tf.
void Packer::make_partm(int cs) {
    const uint64_t n_pieces_total = m_piece_ptr[n_leaves];
    const int ps                  = 1 - cs;
    const uint64_t base           = tf_elems + (cs == 1 ? m_partm_total[0] : 0u);
    m_partm_off[cs].assign(n_pieces_total, kDirect);
    side[cs].partm_base = base;
    uint64_t off        = 0;
    for (int64_t i = 0; i < n_leaves; i++) {
        const htb_leaf &l = m_leaves[i];
        if (!active(l) || (l.rank < 0 && cs == 0)) // dense leaves, direction 0: x rows are read from the input directly
            continue;
        const int n_prod = m_nchunks[ps][i];
        if (n_prod <= 1)
            continue;
        for (int p = 0; p < n_pieces(l); p++) {
            const int len                       = piece_len(l, p);
            m_partm_off[cs][m_piece_ptr[i] + p] = static_cast<uint32_t>(base + off);
            side[cs].combine_m.push_back(CombineEntry{static_cast<uint32_t>(base + off), m_tf_off[m_piece_ptr[i] + p], 0u,
                                                      static_cast<uint32_t>(n_prod) | (static_cast<uint32_t>(len) << 24) | ((l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) ? 0x80000000u : 0u)});
            off += static_cast<uint64_t>(n_prod) * len;
        }
    }
    side[cs].partm_elems = off;
    if (off != m_partm_total[cs])
        throw std::runtime_error("internal: size of a multi-RHS partial area");
}

void Packer::make_mtables() {
    const bool mt_timing = std::getenv("HTB_PACK_TIMING") != nullptr;
    auto mt_last         = std::chrono::steady_clock::now();
    auto mt_lap          = [&](const char *what) {
        if (!mt_timing)
            return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[htb pack]   mtables %-18s %.3f s\n", what, std::chrono::duration<double>(now - mt_last).count());
        mt_last = now;
    };
    const uint64_t base = tf_elems + m_partm_total[0] + m_partm_total[1]; // (make_tf, make_partm: run by the constructor beside the other serial passes)
    mscratch_elems = base;
    if (mscratch_elems >= (uint64_t(1) << 31))
        throw std::runtime_error("multi-RHS scratch exceeds 2^31 vectors");

    std::vector<std::vector<NfSrc>> nf_src_side[2];
    for (int s = 0; s < 2; s++) {
        const int nb = static_cast<int>(side[s].blocks.size());
        side[s].munits.resize(m_unit_ptr[s].back()); // (RawVector: every entry is written by the walk below)
        // the aux records (runs + column tables, store.hpp) go straight to their place: layout_block sized them
        std::vector<std::vector<DenseTask>> tasks(nb), lr_tasks(nb);
        std::vector<std::vector<NfSrc>> &nf_src = nf_src_side[s];
        nf_src.assign(s == 0 && opt.near_field && opt.sort_units ? nb : 0, {});
        bool aux_mismatch = false;
#pragma omp parallel for schedule(dynamic, 64)
        for (int b = 0; b < nb; b++) {
            uint32_t q = 0; // stage of the block being closed
            // same walk and the same cuts as layout_block / fill_block; inside a stage panel units come first
            StageCutter cut{esize, static_cast<uint32_t>(opt.stage_bytes), static_cast<uint32_t>(opt.cseg_bytes)};
            std::vector<UnitSpec> pending;
            std::vector<RunDesc> runs;
            std::vector<uint32_t> col_out, col_src;
            uint64_t pos = m_unit_ptr[s][m_csr_ptr[s][b]];
            auto close   = [&]() {
                if (cut.nu == 0)
                    return;
                std::stable_partition(pending.begin(), pending.end(), [](const UnitSpec &u) { return u.kind != UNIT_ADDVEC; });
                // runs: consecutive panel units on the same rows (and the same applied-twice flag) are one h x K panel
                {
                    runs.clear(), col_out.clear(), col_src.clear(); // (per-block buffers: a stage must not cost three allocations)
                    uint32_t eoff = 0;
                    for (const UnitSpec &u : pending) {
                        if (u.kind == UNIT_ADDVEC)
                            break;
                        const htb_leaf &l = m_leaves[u.leaf];
                        const uint64_t gp = m_piece_ptr[u.leaf] + u.piece;
                        if (runs.empty() || runs.back().row0 != u.row0 || static_cast<uint32_t>(runs.back().h_minus_1) + 1u != u.h || (runs.back().flags & 1u) != u.twice ||
                            static_cast<uint32_t>(runs.back().K) + u.w > 0xffffu)
                            runs.push_back(RunDesc{eoff, static_cast<uint16_t>(col_out.size()), 0, static_cast<uint8_t>(u.row0), static_cast<uint8_t>(u.h - 1u), static_cast<uint8_t>(u.twice), 0, 0, 0xffffu});
                        runs.back().K = static_cast<uint16_t>(runs.back().K + u.w);
                        if (u.kind == UNIT_LOWRANK)
                            runs.back().K_lr = runs.back().K; // (low-rank units come first: a prefix)
                        if (u.kind == UNIT_DENSE && !nf_src.empty()) {
                            const StageDesc &sd = side[s].stages[side[s].blocks[b].first_stage + q];
                            nf_src[b].push_back(NfSrc{sd.byte_off + cut.header_bytes() + static_cast<uint64_t>(eoff) * esize, static_cast<uint32_t>(l.col_offset) + u.k0, static_cast<uint16_t>(u.row0), static_cast<uint16_t>(u.h),
                                                      static_cast<uint16_t>(u.w), static_cast<uint16_t>(u.ld)});
                        }
                        // producer role (this side is streamed by REDUCE_M; consumer side cs = 1 - s)
                        uint32_t out = 0;
                        if (!(l.rank < 0 && s == 1)) { // (dense leaves hold no panel on side 1)
                            const uint32_t po = m_partm_off[1 - s][gp];
                            out               = po == kDirect ? m_tf_off[gp] : po + static_cast<uint32_t>(u.chunk) * static_cast<uint32_t>(piece_len(l, u.piece));
                        }
                        if (u.kind == UNIT_DENSE && !l.data0) {
                            const StageDesc &sd = side[s].stages[side[s].blocks[b].first_stage + q];
                            tasks[b].push_back(DenseTask{sd.byte_off + cut.header_bytes() + static_cast<uint64_t>(eoff) * esize, l.row_offset, l.col_offset, static_cast<int32_t>(u.p0), static_cast<int32_t>(u.k0),
                                                         static_cast<uint16_t>(u.h), static_cast<uint16_t>(u.w), static_cast<uint16_t>(u.ld),
                                                         static_cast<uint16_t>(l.flags & (HTB_LEAF_DIAG_SYMMETRIC | HTB_LEAF_DIAG_HERMITIAN | HTB_LEAF_UPLO_UPPER))});
                        }
                        if (u.kind == UNIT_LOWRANK && !l.data0) { // factors in the device pool: the panel is copied on the device (aca.cu)
                            const StageDesc &sd = side[s].stages[side[s].blocks[b].first_stage + q];
                            lr_tasks[b].push_back(DenseTask{sd.byte_off + cut.header_bytes() + static_cast<uint64_t>(eoff) * esize, static_cast<int32_t>(u.leaf), s, static_cast<int32_t>(u.p0), static_cast<int32_t>(u.k0),
                                                            static_cast<uint16_t>(u.h), static_cast<uint16_t>(u.w), static_cast<uint16_t>(u.ld), 0});
                        }
                        for (uint32_t k = 0; k < u.w; k++) {
                            col_out.push_back(out + k);
                            col_src.push_back(u.kind == UNIT_LOWRANK ? m_tf_off[gp] + k : (0x80000000u | (static_cast<uint32_t>(l.col_offset) + u.k0 + k)));
                        }
                        eoff += u.elems();
                    }
                    const size_t ncols4 = (col_out.size() + 3u) & ~size_t(3);
                    const size_t bytes  = sizeof(AuxHeader) + runs.size() * sizeof(RunDesc) + ncols4 * 4u;
                    if (q >= side[s].blocks[b].n_stages) {
                        aux_mismatch = true; // (reported after the parallel loop)
                        return;
                    }
                    const StageDesc &sd_q = side[s].stages[side[s].blocks[b].first_stage + q];
                    if (bytes != static_cast<size_t>(sd_q.flags >> 1) * 16u) {
                        aux_mismatch = true; // (nothing is written past a record)
                        return;
                    }
                    col_out.resize(ncols4, 0u);
                    col_src.resize(ncols4, 0u);
                    const AuxHeader ah{static_cast<uint32_t>(runs.size()), static_cast<uint32_t>(col_out.size()), {0u, 0u}};
                    for (int role = 0; role < 2; role++) {
                        unsigned char *dst = (role == 0 ? side[s].aux_reduce : side[s].aux_apply).data() + static_cast<size_t>(sd_q.aux_off16) * 16u;
                        std::memcpy(dst, &ah, sizeof(ah));
                        if (!runs.empty())
                            std::memcpy(dst + sizeof(ah), runs.data(), runs.size() * sizeof(RunDesc));
                        if (ncols4)
                            std::memcpy(dst + sizeof(ah) + runs.size() * sizeof(RunDesc), (role == 0 ? col_out : col_src).data(), ncols4 * 4u);
                    }
                    q++;
                }
                uint32_t poff = 0, in_batch = 0;
                bool first    = true;
                for (const UnitSpec &u : pending) {
                    const htb_leaf &l = m_leaves[u.leaf];
                    const uint64_t gp = m_piece_ptr[u.leaf] + u.piece;
                    MUnit mu{0, 0, 0, 0};
                    // producer role: this unit's side is ps, the consumer side is 1 - s
                    const int cs = 1 - s;
                    if (u.kind != UNIT_ADDVEC && !(l.rank < 0 && cs == 0)) {
                        const uint32_t po = m_partm_off[cs][gp];
                        mu.out            = po == kDirect ? m_tf_off[gp] : po + static_cast<uint32_t>(u.chunk) * static_cast<uint32_t>(piece_len(l, u.piece));
                    }
                    // consumer role: this unit's side is the consumer side
                    mu.flags = (u.kind == UNIT_ADDVEC ? u.h : u.w) << 8;
                    if (u.kind == UNIT_LOWRANK) {
                        mu.src = m_tf_off[gp];
                        mu.flags |= 2u;
                    }
                    else if (u.kind == UNIT_DENSE)
                        mu.src = 0x80000000u | (static_cast<uint32_t>(l.col_offset) + u.k0);
                    else
                        mu.src = m_tf_off[gp] + u.sub_off;
                    if (u.kind != UNIT_ADDVEC) {
                        const uint32_t need = munit_ld(u.row0, u.h) * u.w;
                        if (first || poff + need > kPanelBufferElems || in_batch == 32) { // <= 32 units per batch: one lane each
                            mu.flags |= 1u;
                            poff     = 0;
                            in_batch = 0;
                            first    = false;
                        }
                        mu.poff = poff;
                        poff += need;
                        in_batch++;
                    }
                    side[s].munits[pos++] = mu;
                }
                cut.reset();
                pending.clear();
            };
            walk_block(s, b, [&](const UnitSpec &u) {
                if (!cut.fits(u.elems(), u.celems()))
                    close();
                cut.add(u.elems(), u.celems());
                pending.push_back(u);
            });
            close();
            if (q != side[s].blocks[b].n_stages)
                aux_mismatch = true;
        }
        mt_lap(s == 0 ? "walk side 0" : "walk side 1");
        if (aux_mismatch)
            throw std::runtime_error("internal: aux records do not match the stages");
        // concatenate the per-block task lists (offsets by prefix sums, copies in parallel)
        std::vector<uint64_t> dt_at(static_cast<size_t>(nb) + 1, 0), lt_at(static_cast<size_t>(nb) + 1, 0);
        for (int b = 0; b < nb; b++) {
            dt_at[b + 1] = dt_at[b] + tasks[b].size();
            lt_at[b + 1] = lt_at[b] + lr_tasks[b].size();
        }
        side[s].dense_tasks.resize(dt_at[nb]);
        side[s].lr_tasks.resize(lt_at[nb]);
#pragma omp parallel for schedule(dynamic, 64)
        for (int b = 0; b < nb; b++) {
            if (!tasks[b].empty())
                std::memcpy(static_cast<void *>(side[s].dense_tasks.data() + dt_at[b]), tasks[b].data(), tasks[b].size() * sizeof(DenseTask));
            if (!lr_tasks[b].empty())
                std::memcpy(static_cast<void *>(side[s].lr_tasks.data() + lt_at[b]), lr_tasks[b].data(), lr_tasks[b].size() * sizeof(DenseTask));
        }
        mt_lap(s == 0 ? "concat side 0" : "concat side 1");
    }
    if (!nf_src_side[0].empty())
        make_near_field(nf_src_side[0]); // (after both sides: the aux slot size is known)
}

// The multi-RHS near field of side 0 (store.hpp, NearFieldLayout): per target block, the union of the columns of its dense
// units, one full-height column-major panel cut into stages, and the copy tasks that fill it from the main stream.
void Packer::make_near_field(const std::vector<std::vector<NfSrc>> &per_block) {
    const int nb = static_cast<int>(per_block.size());
    struct PerBlock {
        std::vector<StageDesc> stages;
        std::vector<unsigned char> aux;
        std::vector<NfTask> tasks;
        uint64_t bytes = 0, coefficients = 0;
    };
    std::vector<PerBlock> pb(nb);
    // a stage's aux record [AuxHeader | RunDesc | K columns] must fit the aux part of the ring slots, sized by the main sides
    const uint32_t aux_cap = std::max<uint32_t>(64u, std::max(side[0].aux_max_bytes, side[1].aux_max_bytes));
    const uint32_t k_cap   = std::max<uint32_t>(4u, std::min<uint32_t>(512u, ((aux_cap - 32u) / 4u) & ~3u));
    const uint32_t sub_rows = opt.nf_rows > 0 ? static_cast<uint32_t>((opt.nf_rows + 7) & ~7) : 0u; // rows of a sub-panel (0: the whole block)
#pragma omp parallel for schedule(dynamic, 64)
    for (int b = 0; b < nb; b++) {
        const std::vector<NfSrc> &src = per_block[b];
        if (src.empty())
            continue;
        const uint32_t nrows = static_cast<uint32_t>(side[0].blocks[b].nrows);
        PerBlock &out        = pb[b];
        // One panel per group of `sub_rows` rows of the block (tile-aligned): the fewer leaf clusters share a panel, the fewer
        // zeros it holds (41 % of the entries of a whole-block panel are coefficients at N = 1e6) — and the more often a B row is fetched.
        for (uint32_t r0 = 0; r0 < nrows; r0 += (sub_rows ? sub_rows : nrows)) {
            const uint32_t r1 = sub_rows ? std::min(nrows, r0 + sub_rows) : nrows, hsub = r1 - r0;
            std::vector<uint32_t> cols;
            for (const NfSrc &u : src)
                if (u.row0 < r1 && static_cast<uint32_t>(u.row0) + u.h > r0)
                    for (uint32_t k = 0; k < u.w; k++)
                        cols.push_back(u.col + k);
            if (cols.empty())
                continue;
            std::sort(cols.begin(), cols.end());
            cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
            const uint32_t ld  = unit_ld(hsub, esize);
            const uint32_t kst = std::max<uint32_t>(1u, std::min<uint32_t>(k_cap, static_cast<uint32_t>((static_cast<size_t>(opt.stage_bytes) - 16u) / (static_cast<size_t>(ld) * esize))));
            std::vector<uint64_t> stage_off;
            std::vector<size_t> run_at; // where the stage's RunDesc sits in out.aux
            for (uint32_t c0 = 0; c0 < cols.size(); c0 += kst) {
                const uint32_t K = std::min<uint32_t>(kst, static_cast<uint32_t>(cols.size()) - c0), K4 = (K + 3u) & ~3u;
                StageDesc sd{};
                sd.byte_off = out.bytes;
                sd.nbytes   = round16(16u + static_cast<uint32_t>(static_cast<size_t>(K) * ld * esize));
                const uint32_t aux_bytes = static_cast<uint32_t>(sizeof(AuxHeader) + sizeof(RunDesc)) + 4u * K4;
                sd.flags     = static_cast<uint16_t>((aux_bytes / 16u) << 1);
                sd.aux_off16 = static_cast<uint32_t>(out.aux.size() / 16u);
                stage_off.push_back(out.bytes);
                out.bytes += sd.nbytes;
                out.coefficients += static_cast<uint64_t>(K) * hsub;
                out.stages.push_back(sd);
                const AuxHeader ah{1u, K4, {0u, 0u}};
                const RunDesc rd{0u, 0, static_cast<uint16_t>(K), static_cast<uint8_t>(r0), static_cast<uint8_t>(hsub - 1u), 0, 0, 0, 0};
                const size_t at = out.aux.size();
                run_at.push_back(at + sizeof(AuxHeader));
                out.aux.resize(at + aux_bytes, 0);
                std::memcpy(out.aux.data() + at, &ah, sizeof(ah));
                std::memcpy(out.aux.data() + at + sizeof(ah), &rd, sizeof(rd));
                uint32_t *ct = reinterpret_cast<uint32_t *>(out.aux.data() + at + sizeof(ah) + sizeof(rd));
                for (uint32_t k = 0; k < K; k++)
                    ct[k] = 0x80000000u | cols[c0 + k];
            }
            for (const NfSrc &u : src) {
                const uint32_t i0 = std::max<uint32_t>(u.row0, r0), i1 = std::min<uint32_t>(static_cast<uint32_t>(u.row0) + u.h, r1);
                if (i0 >= i1)
                    continue;
                const uint32_t p = static_cast<uint32_t>(std::lower_bound(cols.begin(), cols.end(), u.col) - cols.begin()); // (consecutive columns stay consecutive in the union)
                for (uint32_t done = 0; done < u.w;) {
                    const uint32_t st = (p + done) / kst, in_stage = (p + done) - st * kst;
                    const uint32_t w  = std::min<uint32_t>(u.w - done, kst - in_stage);
                    out.tasks.push_back(NfTask{u.src_off + (static_cast<uint64_t>(done) * u.ld + (i0 - u.row0)) * esize, stage_off[st] + 16u + (static_cast<uint64_t>(in_stage) * ld + (i0 - r0)) * esize,
                                               static_cast<uint16_t>(i1 - i0), static_cast<uint16_t>(w), u.ld, static_cast<uint16_t>(ld), {0u, 0u}});
                    RunDesc *rd = reinterpret_cast<RunDesc *>(out.aux.data() + run_at[st]); // row tiles of the block (8 real rows) the stage has coefficients in
                    const uint32_t rows_per_tile = 8u; // (APPLY_M tiles are 8 rows of the block for both coefficient types)
                    for (uint32_t t = i0 / rows_per_tile; t <= (i1 - 1u) / rows_per_tile && t < 16u; t++)
                        rd->tiles = static_cast<uint16_t>(rd->tiles | (1u << t));
                    done += w;
                }
            }
        }
    }
    // concatenate (block order); byte / aux offsets become side-wide
    nf = NearFieldLayout{};
    uint64_t stream = 0, aux = 0;
    for (int b = 0; b < nb; b++) {
        PerBlock &in = pb[b];
        if (in.stages.empty())
            continue;
        BlockDesc bd       = side[0].blocks[b];
        bd.first_stage     = static_cast<uint32_t>(nf.stages.size());
        bd.n_stages        = static_cast<uint32_t>(in.stages.size());
        bd.flags           = 0;
        bd.n_twice_stages  = 0;
        nf.blocks.push_back(bd);
        for (StageDesc sd : in.stages) {
            sd.byte_off += stream;
            sd.aux_off16 += static_cast<uint32_t>(aux / 16u);
            nf.aux_max_bytes = std::max<uint32_t>(nf.aux_max_bytes, static_cast<uint32_t>(sd.flags >> 1) * 16u);
            nf.stages.push_back(sd);
        }
        for (NfTask t : in.tasks) {
            t.dst_off += stream;
            nf.tasks.push_back(t);
        }
        nf.aux_apply.insert(nf.aux_apply.end(), in.aux.begin(), in.aux.end());
        stream += in.bytes;
        aux += in.aux.size();
        nf.coefficients += in.coefficients;
        PerBlock().tasks.swap(in.tasks);
    }
    if (aux / 16u >= (uint64_t(1) << 32)) {
        nf = NearFieldLayout{};
        return;
    }
    nf.stream_bytes = stream;
    const StageHeader hdr{0u, 16u, 0u, 0u};
    nf.headers.resize(nf.stages.size() * sizeof(StageHeader));
    nf.hdr_off.resize(nf.stages.size() + 1);
    for (size_t st = 0; st < nf.stages.size(); st++) {
        std::memcpy(nf.headers.data() + st * sizeof(StageHeader), &hdr, sizeof(hdr));
        nf.hdr_off[st] = st * sizeof(StageHeader);
    }
    nf.hdr_off[nf.stages.size()] = nf.stages.size() * sizeof(StageHeader);
    if (std::getenv("HTB_PACK_TIMING"))
        std::fprintf(stderr, "[htb pack] near field: %zu blocks, %zu stages, %.1f M panel entries (%.2f GB), %zu copy tasks, aux %.1f MB\n", nf.blocks.size(), nf.stages.size(), nf.coefficients / 1e6, nf.stream_bytes / 1e9,
                     nf.tasks.size(), nf.aux_apply.size() / 1e6);
    nf.order.resize(nf.blocks.size());
    std::iota(nf.order.begin(), nf.order.end(), 0u);
    std::stable_sort(nf.order.begin(), nf.order.end(), [&](uint32_t a, uint32_t c) { return nf.blocks[a].n_stages > nf.blocks[c].n_stages; });
}

// Where the REDUCE result of a producer unit of side ps goes.
uint32_t Packer::producer_out(int ps, const UnitSpec &u) const {
    const int cs      = 1 - ps;
    const htb_leaf &l = m_leaves[u.leaf];
    const uint32_t po = m_part_off[cs][m_piece_ptr[u.leaf] + u.piece];
    if (po != kDirect) {
        if (u.kind == UNIT_ADDVEC)
            return po + u.sub_off; // copy kind: every producer fills its sub-range of the piece
        return po + static_cast<uint32_t>(u.chunk) * static_cast<uint32_t>(piece_len(l, u.piece));
    }
    // direct: the single consumer unit of the piece
    if (l.rank < 0 && cs == 1) {
        const int c = m_blk_of[1][l.col_offset + u.piece * piece] - m_first_blk[1][u.leaf];
        int p_lo, p_hi;
        addvec_pieces(u.leaf, c, p_lo, p_hi);
        return m_unit_slot[1][unit_of(1, u.leaf, c, u.piece - p_lo)];
    }
    return m_unit_slot[cs][unit_of(cs, u.leaf, 0, u.piece)];
}

template <typename Emit>
void Packer::walk_block(int s, int b, Emit &&emit) const {
    const int bs = m_block_start[s][b];
    const uint64_t e_end = m_csr_ptr[s][b + 1];
    for (uint64_t e = m_csr_ptr[s][b]; e < e_end; e++) {
        if (e + 8 < e_end)
            prefetch_leaf(s, m_csr_leaf[s][e + 8]); // (the leaves of a block are scattered over a list of millions)
        const uint32_t li = m_csr_leaf[s][e];
        const htb_leaf &l = m_leaves[li];
        const int a       = start_of(s, l);
        const int chunk   = b - m_first_blk[s][li];
        int lo, hi;
        chunk_range(s, li, chunk, lo, hi);
        UnitSpec u{};
        u.leaf  = li;
        u.chunk = chunk;
        u.twice = (l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) ? 1u : 0u;
        uint64_t ui = m_unit_ptr[s][e];
        if (l.rank < 0 && s == 1) {
            int p_lo, p_hi;
            addvec_pieces(li, chunk, p_lo, p_hi);
            for (int p = p_lo; p <= p_hi; p++) {
                const int sub_lo = std::max(lo, p * piece), sub_hi = std::min(hi, p * piece + piece_len(l, p));
                u.ui      = ui++;
                u.piece   = p;
                u.kind    = UNIT_ADDVEC;
                u.row0    = static_cast<uint32_t>(a + sub_lo - bs);
                u.h       = static_cast<uint32_t>(sub_hi - sub_lo);
                u.p0      = static_cast<uint32_t>(sub_lo);
                u.k0      = 0;
                u.w       = 0;
                u.sub_off = static_cast<uint32_t>(sub_lo - p * piece);
                u.ld      = 0;
                emit(u);
            }
            continue;
        }
        u.kind = l.rank < 0 ? UNIT_DENSE : UNIT_LOWRANK;
        u.row0 = static_cast<uint32_t>(a + lo - bs);
        u.h    = static_cast<uint32_t>(hi - lo);
        u.p0   = static_cast<uint32_t>(lo);
        u.ld   = unit_ld(u.h, esize);
        for (int p = 0; p < n_pieces(l); p++) {
            u.ui    = ui++;
            u.piece = p;
            u.k0    = static_cast<uint32_t>(p * piece);
            u.w     = static_cast<uint32_t>(piece_len(l, p));
            emit(u);
        }
    }
}

void Packer::layout_block(int s, int b, std::vector<StageDesc> &stages, RawVector<uint32_t> &unit_stage, uint64_t &n_units, bool &any_twice) {
    StageCutter cut{esize, static_cast<uint32_t>(opt.stage_bytes), static_cast<uint32_t>(opt.cseg_bytes)};
    uint64_t off     = 0;
    uint16_t stflags = 0;
    uint32_t first   = 0; // units of the block before the current stage
    uint32_t n_panel = 0; // coefficient-carrying units of the current stage
    // size of the stage's multi-RHS aux record (make_mtables writes it: same run detection): it travels in aux_off16 until the
    // constructor turns the sizes into offsets
    uint32_t n_runs = 0, n_cols = 0, run_row0 = 0, run_h = 0, run_twice = 0, run_K = 0;
    auto close       = [&]() {
        if (cut.nu == 0)
            return;
        const uint32_t aux_bytes = static_cast<uint32_t>(sizeof(AuxHeader) + n_runs * sizeof(RunDesc) + ((n_cols + 3u) & ~3u) * 4u);
        stages.push_back(StageDesc{off, cut.nbytes(), 0u, static_cast<uint16_t>(cut.c_len_padded()), stflags, first, static_cast<uint16_t>(cut.nu), static_cast<uint16_t>(n_panel), aux_bytes});
        n_panel = 0;
        n_runs = n_cols = 0;
        off += cut.nbytes();
        first += cut.nu;
        cut.reset();
        stflags = 0;
    };
    walk_block(s, b, [&](const UnitSpec &u) {
        if (!cut.fits(u.elems(), u.celems()))
            close();
        unit_stage[u.ui]      = static_cast<uint32_t>(stages.size());
        m_unit_cslot[s][u.ui] = static_cast<uint16_t>(cut.c_elems);
        cut.add(u.elems(), u.celems());
        n_panel += u.kind != UNIT_ADDVEC ? 1u : 0u;
        if (u.kind != UNIT_ADDVEC) {
            if (n_runs == 0 || run_row0 != u.row0 || run_h != u.h || run_twice != u.twice || run_K + u.w > 0xffffu) {
                n_runs++;
                run_row0 = u.row0, run_h = u.h, run_twice = u.twice, run_K = 0;
            }
            run_K += u.w;
            n_cols += u.w;
        }
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
void Packer::fill_block(int s, int b, char *dst, bool headers_only) const {
    StageCutter cut{esize, static_cast<uint32_t>(opt.stage_bytes), static_cast<uint32_t>(opt.cseg_bytes)};
    std::vector<UnitSpec> pending;
    char *cursor        = dst;
    uint64_t first_unit = m_unit_ptr[s][m_csr_ptr[s][b]]; // index of the stage's first unit in the side's MUnit table
    auto close          = [&]() {
        if (cut.nu == 0)
            return;
        const uint32_t nbytes = cut.nbytes();
        // coefficient-carrying units first, ADDVEC units last (the kernels walk the two groups differently)
        std::stable_partition(pending.begin(), pending.end(), [](const UnitSpec &u) { return u.kind != UNIT_ADDVEC; });
        const uint32_t n_panel = static_cast<uint32_t>(std::count_if(pending.begin(), pending.end(), [](const UnitSpec &u) { return u.kind != UNIT_ADDVEC; }));
        StageHeader hdr{cut.nu, cut.header_bytes(), n_panel, static_cast<uint32_t>(first_unit)};
        first_unit += cut.nu;
        std::memcpy(cursor, &hdr, sizeof(hdr));
        Unit *units   = reinterpret_cast<Unit *>(cursor + sizeof(StageHeader));
        T *data       = reinterpret_cast<T *>(cursor + cut.header_bytes());
        uint32_t eoff = 0;
        for (uint32_t i = 0; i < cut.nu; i++) {
            const UnitSpec &u = pending[i];
            units[i]          = Unit{eoff, make_geom(u.row0, u.h, u.w, u.kind, u.twice), producer_out(s, u), m_unit_cslot[s][u.ui], 0};
            if (u.kind == UNIT_ADDVEC)
                continue;
            if (headers_only) { // (the panels are filled on the device)
                eoff += u.elems();
                continue;
            }
            const htb_leaf &l = m_leaves[u.leaf];
            T *out            = data + eoff;
            if (u.kind == UNIT_LOWRANK && !l.data0) {
                // factors in the device pool (htb_create_compressed): the panel travels as zeros and is copied on the device
                for (uint32_t k = 0; k < u.w; k++)
                    std::memset(static_cast<void *>(out + static_cast<size_t>(k) * u.ld), 0, sizeof(T) * u.h);
            } else if (u.kind == UNIT_LOWRANK && s == 1) {
                // Vt panel: element (i, k) = V[k + (p0+i) * r], V is r x n column-major
                const T *V     = static_cast<const T *>(l.data1);
                const size_t r = static_cast<size_t>(l.rank);
                for (uint32_t k = 0; k < u.w; k++)
                    for (uint32_t i = 0; i < u.h; i++)
                        out[i + static_cast<size_t>(k) * u.ld] = V[(u.k0 + k) + (u.p0 + i) * r];
            } else if (u.kind == UNIT_DENSE && !l.data0) {
                // generated on the device after the upload (htb_create_generated): the panel travels as zeros
                for (uint32_t k = 0; k < u.w; k++)
                    std::memset(static_cast<void *>(out + static_cast<size_t>(k) * u.ld), 0, sizeof(T) * u.h);
            } else if (u.kind == UNIT_DENSE && (l.flags & (HTB_LEAF_DIAG_SYMMETRIC | HTB_LEAF_DIAG_HERMITIAN))) {
                // symv / hemv read only the UPLO triangle (add_matrix_vector_product.hpp:26-52): rebuild the full
                // block from that triangle so the kernels see an ordinary dense leaf
                const T *A       = static_cast<const T *>(l.data0);
                const size_t m   = static_cast<size_t>(l.nb_rows);
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
                        out[i + static_cast<size_t>(k) * u.ld] = v;
                    }
            } else {
                // U panel or dense leaf: column-major with lda = nb_rows
                const T *A     = static_cast<const T *>(l.data0);
                const size_t m = static_cast<size_t>(l.nb_rows);
                for (uint32_t k = 0; k < u.w; k++)
                    std::memcpy(out + static_cast<size_t>(k) * u.ld, A + u.p0 + (u.k0 + k) * m, sizeof(T) * u.h);
            }
            if (u.ld != u.h) // zero pad rows
                for (uint32_t k = 0; k < u.w; k++)
                    for (uint32_t i = u.h; i < u.ld; i++)
                        out[i + static_cast<size_t>(k) * u.ld] = T(0);
            eoff += u.elems();
        }
        const size_t used = cut.header_bytes() + static_cast<size_t>(cut.data_elems) * esize;
        if (headers_only)
            cursor += cut.header_bytes();
        else {
            if (used < nbytes)
                std::memset(cursor + used, 0, nbytes - used);
            cursor += nbytes;
        }
        cut.reset();
        pending.clear();
    };
    walk_block(s, b, [&](const UnitSpec &u) {
        if (!cut.fits(u.elems(), u.celems()))
            close();
        cut.add(u.elems(), u.celems());
        pending.push_back(u);
    });
    close();
}

void Packer::fill_headers(int s, int b0, int b1, char *dst) const {
    const uint64_t base = m_hdr_off[s][side[s].blocks[b0].first_stage];
#pragma omp parallel for schedule(dynamic, 16)
    for (int b = b0; b < b1; b++) {
        char *p = dst + (m_hdr_off[s][side[s].blocks[b].first_stage] - base);
        if (dtype == HTB_DOUBLE)
            fill_block<double>(s, b, p, true);
        else
            fill_block<std::complex<double>>(s, b, p, true);
    }
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

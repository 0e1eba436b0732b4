// htool_b200/csrc/store.hpp — the device-resident flattened leaf store: formats shared by the host
// packer (packer.cpp) and the sm_100a kernels (kernels.cu).
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
//     CHUNKS (h x W, h <= block_rows) and along its short dimension into PIECES of <= piece_cols columns
//     (the same cut on both sides of a leaf); one (chunk, piece) is a UNIT (h x w), stored column-major
//     with leading dimension ld = h rounded up to an even number for double (so that every column starts on a
//     16 B boundary and a lane can fetch two rows with one 128-bit shared-memory load; pad rows are zero) and
//     ld = h for complex<double>. Rank and size bucketing happen through the (h, w) unit shape;
//   * all units of one block are concatenated, in leaf order, into the block's STREAM, itself cut into
//     STAGES (<= stage_bytes, 16 B aligned) = one bulk-async copy each. A stage starts with its own
//     unit descriptors, so a CTA needs nothing but the stream: [StageHeader | Unit x n | coefficients];
//   * every coefficient is stored exactly once per side it is needed on and read exactly once per pass.
//
// Two generic passes run over a side (kernels.cu):
//   REDUCE  partial[k] = sum_i op(P[i,k]) * in[i]     (t = V x, or t = op(U)^T x / z = op(A)^T x)
//   APPLY   out[i]    += sum_k op(P[i,k]) * c[k]      (U t, A x, or op(Vt) t; + ADDVEC z)
// and every product of the reference (N / T / C, plain / symmetric-twice) is a fixed sequence of them.
//
// How the c vectors reach APPLY (the "c-stream"): the scratch holds, for each side s, an array CS[s] with one
// slot per unit of side s, IN THE ORDER OF THE SIDE'S STREAM, so that the slots of one stage are one
// contiguous segment which the APPLY producer bulk-copies into shared memory next to the stage itself.
// The slots are filled by the pass over the OTHER side:
//   direction 0 (consumer side 0: y = A x):      REDUCE over side 1 writes t pieces (low rank) and x slices
//                                                (dense leaves, through the coefficient-less ADDVEC units);
//   direction 1 (consumer side 1: y = op(A)^T x): REDUCE over side 0 writes t' = op(U)^T x pieces and
//                                                z = op(A)^T x slices (consumed by the ADDVEC units).
// When a piece has exactly one producer unit and one consumer unit the producer writes straight into the
// consumer's slot; otherwise producers write partials (PART[dir]) and a COMBINE pass sums them in chunk
// order and replicates the result into every consumer slot.
// One scratch copy = [PART[0] | PART[1] | CS[0] | CS[1]]; all offsets below are element offsets into a copy.
#ifndef HTB_STORE_HPP
#define HTB_STORE_HPP

#include <cstddef>
#include <cstdint>
#include <memory>
#include <new>
#include <type_traits>
#include <cstring>
#include <utility>
#include <vector>

#ifdef __CUDACC__
#    define HTB_HD __host__ __device__
#else
#    define HTB_HD
#endif

namespace htb {

enum UnitKind : uint32_t { UNIT_LOWRANK = 0, // panel of U (side 0) or of Vt (side 1)
                           UNIT_DENSE   = 1, // panel of a dense leaf (side 0 only)
                           UNIT_ADDVEC  = 2  // side 1 only, no coefficients. REDUCE: slot <- in[rows] (x slice of a dense leaf).
                                             //                               APPLY:  out[rows] += slot (z of a dense leaf applied transposed)
};

// geom: row0 [0:8) | h-1 [8:16) | w [16:24) | kind [24:26) | applied_twice [26]
struct Unit {
    uint32_t data_off; // element offset of the h x w panel inside the stage's coefficient region
    uint32_t geom;
    uint32_t out;   // REDUCE: scratch offset receiving this unit's result (w values; ADDVEC: h values)
    uint16_t cslot; // APPLY: element offset of this unit's c vector inside the stage's c segment (w values; ADDVEC: h values)
    uint16_t reserved;
};
static_assert(sizeof(Unit) == 16, "Unit must be 16 bytes");

HTB_HD inline uint32_t unit_row0(uint32_t g) { return g & 0xffu; }
HTB_HD inline uint32_t unit_h(uint32_t g) { return ((g >> 8) & 0xffu) + 1u; }
HTB_HD inline uint32_t unit_w(uint32_t g) { return (g >> 16) & 0xffu; }
HTB_HD inline uint32_t unit_kind(uint32_t g) { return (g >> 24) & 0x3u; }
HTB_HD inline uint32_t unit_twice(uint32_t g) { return (g >> 26) & 0x1u; }
// leading dimension of a unit of height h (double: even, so that panel columns are 16-byte aligned). A leading dimension = 4 (mod 8)
// doubles for the tall panels (conflict-free DMMA fragment loads: ld = 122 costs two shared-memory wavefronts per load) was measured
// at N = 1e6: mu = 64 18.07 vs 18.28 ms, mu = 1 3.36 vs 3.30 ms — the fragment loads are not what limits the multi-RHS kernels, and the
// 1.2 % of extra bytes cost the single-RHS product more. Not kept.
HTB_HD inline uint32_t unit_ld(uint32_t h, size_t esize) { return esize == 8 ? (h + 1u) & ~1u : h; }
inline uint32_t make_geom(uint32_t row0, uint32_t h, uint32_t w, uint32_t kind, uint32_t twice) {
    return (row0 & 0xffu) | (((h - 1u) & 0xffu) << 8) | ((w & 0xffu) << 16) | ((kind & 3u) << 24) | ((twice & 1u) << 26);
}

struct StageHeader {
    uint32_t n_units;
    uint32_t data_byte_off; // from the start of the stage to the coefficient region (multiple of 16)
    uint32_t n_panel;       // units [0, n_panel) hold coefficients; [n_panel, n_units) are ADDVEC units
    uint32_t first_unit;    // index of the stage's first unit in the side's MUnit table (multi-RHS kernels)
};
static_assert(sizeof(StageHeader) == 16, "StageHeader must be 16 bytes");

struct StageDesc {
    uint64_t byte_off; // into the side's stream, multiple of 16
    uint32_t nbytes;   // multiple of 16
    uint32_t c_off;    // first element of the stage's c segment in CS[side] (segment start is 16 B aligned)
    uint16_t c_len;    // elements of the c segment, padded so that its byte size is a multiple of 16
    uint16_t flags;    // bit 0: holds at least one applied-twice unit; bits [1, 16): length of the stage's multi-RHS aux record in 16 B units
    uint32_t first_unit; // index of the stage's first unit in the side's MUnit table
    uint16_t n_units;    // units of the stage
    uint16_t n_panel;    // of which coefficient-carrying (they come first)
    uint32_t aux_off16;  // offset of the stage's multi-RHS aux record in the side's aux arrays, in 16 B units
};
static_assert(sizeof(StageDesc) == 32, "StageDesc must be 32 bytes");

struct BlockDesc {
    int32_t row_start; // first index of the block in the side's index space
    int32_t nrows;
    uint32_t first_stage;
    uint32_t n_stages;
    uint32_t flags;          // bit 0: holds at least one applied-twice unit
    uint32_t n_twice_stages; // stages holding at least one applied-twice unit
    uint32_t reserved[2];
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc must be 32 bytes");

// One piece that could not be written producer -> consumer directly: v[0..len) = sum_{j < n_sum} scratch[src + j*len ..],
// then v is copied (whole or a sub-range) into each of its n_dst consumer slots.
struct CombineEntry {
    uint32_t src;
    uint32_t dst_first; // index of the first CombineDst
    uint32_t n_dst;
    uint32_t packed; // n_sum [0:24) | len [24:31) | applied_twice [31]
};
static_assert(sizeof(CombineEntry) == 16, "CombineEntry must be 16 bytes");
HTB_HD inline uint32_t combine_n_sum(uint32_t p) { return p & 0xffffffu; }
HTB_HD inline uint32_t combine_len(uint32_t p) { return (p >> 24) & 0x7fu; }
HTB_HD inline uint32_t combine_twice(uint32_t p) { return p >> 31; }

struct CombineDst {
    uint32_t slot;    // scratch offset of the consumer's slot
    uint16_t sub_off; // the consumer wants v[sub_off .. sub_off + sub_len)
    uint16_t sub_len;
};
static_assert(sizeof(CombineDst) == 8, "CombineDst must be 8 bytes");

// ---- multi-RHS (mu >= 8, double) side tables ---------------------------------------------------------------------
// The multi-RHS kernels (mkernels.cu) work on groups of MC <= 64 right-hand sides. Their vectors are MC wide, so the
// c-stream (one replicated slot per consumer unit) would cost MC times its single-RHS size; instead the first pass
// writes each piece ONCE into TF[piece] (or per-chunk partials + COMBINE_M when a piece has several producer chunks)
// and the second pass reads TF / the input matrix rows directly. Offsets are in VECTORS (multiply by MC).
// One MUnit per unit, in the order the units appear in the stages (StageDesc.first_unit + index in the stage).
struct MUnit {
    uint32_t out;  // REDUCE_M: TF / PARTM offset receiving the unit's w result vectors (unused for ADDVEC)
    uint32_t src;  // APPLY_M: TF offset of the unit's w (ADDVEC: h) input vectors; bit 31 set: row index of the INPUT matrix (dense leaf, direction 0)
    uint32_t poff; // APPLY_M: element offset of the unit's padded copy inside the CTA's panel buffer (see mkernels.cu)
    uint32_t flags; // bit 0: this unit starts a new panel-buffer batch; bit 1: low-rank panel (its input vectors live in TF);
                    // bits [8, 16): w (ADDVEC: h), the number of input vectors
};
static_assert(sizeof(MUnit) == 16, "MUnit must be 16 bytes");
// padded panel-buffer geometry of a unit: rows are shifted by row0 & 7 so that 8-row tiles coincide with the
// block's 8-row tiles, and the leading dimension is = 4 (mod 8) doubles: conflict-free DMMA fragment loads
HTB_HD inline uint32_t munit_rows8(uint32_t row0, uint32_t h) { return (((row0 & 7u) + h + 7u) >> 3) << 3; }
HTB_HD inline uint32_t munit_ld(uint32_t row0, uint32_t h) { return munit_rows8(row0, h) + 4u; }
constexpr uint32_t kPanelBufferElems = 4608; // 36 KiB of doubles

// ---- multi-RHS aux records: RUNS ----------------------------------------------------------------------------------
// Inside a block the packer orders the units by (first row, height): all the panels that act on the same rows of the
// block (the U panels / dense leaves of one target cluster, or the V^T panels of one source cluster) are then
// CONSECUTIVE in the stream and, having the same leading dimension, form ONE column-major panel h x K (K = sum of the
// unit widths): a RUN. The multi-RHS kernels contract whole runs — C[rows] += Panel (h x K) . B (K x mu) with K in the
// hundreds — instead of unit by unit, so that no DMMA tile is padded along K (the rank of a leaf is not a multiple of the
// 8 / 4 of m8n8k4) and the per-unit overhead disappears. What a column of a run multiplies / produces is listed per
// COLUMN in the stage's aux record, which the producer lane bulk-copies next to the stage:
//   [AuxHeader | RunDesc x n_runs | uint32 x n_cols (padded to a multiple of 4)]
// Two arrays with identical offsets exist per side: aux_reduce (column entry = scratch vector receiving the column's
// result in REDUCE_M) and aux_apply (column entry = scratch vector, or input-matrix row | bit 31, the column multiplies
// in APPLY_M).
struct AuxHeader {
    uint32_t n_runs, n_cols, reserved[2];
};
static_assert(sizeof(AuxHeader) == 16, "AuxHeader must be 16 bytes");
struct RunDesc {
    uint32_t data_off; // element offset of the run's panel inside the stage's coefficient region
    uint16_t col0;     // first column of the run in the stage's column table
    uint16_t K;        // columns of the run
    uint8_t row0;      // first row inside the block
    uint8_t h_minus_1;
    uint8_t flags;     // bit 0: applied-twice units (runs are homogeneous)
    uint8_t reserved8;
    uint16_t K_lr;     // the first K_lr columns of the run belong to low-rank units, the dense ones follow (sorted so by the packer)
    uint16_t tiles;    // bit t set = the run has coefficients in rows [8 t, 8 t + 8) of the block (0xffff for the runs of the main streams;
                       // near-field panels are block-sparse. Skipping the empty tiles in APPLY_M was measured: no gain)
};
static_assert(sizeof(RunDesc) == 16, "RunDesc must be 16 bytes");
// capacity of a stage's aux record: a stage holds <= cseg_bytes / esize columns (c-segment limit), one run per column at worst
HTB_HD inline uint32_t aux_slot_bytes(uint32_t cseg_bytes) { return (16u + 20u * (cseg_bytes / 8u) + 127u) & ~127u; }

// ---- leaf assembly on the device: one task per dense unit whose leaf came without host data ------------
// The generate kernel (generate.cu) evaluates the built-in kernel function at the unit's points straight into the
// uploaded stream: coefficient (i, k) of the unit is entry (p0 + i, k0 + k) of the leaf whose first row / column are
// lrow / lcol in the root block's cluster numbering. flags = the leaf's HTB_LEAF_DIAG_* / UPLO bits (symv / hemv leaves
// are rebuilt from their UPLO triangle exactly as the packer does for host data).
struct DenseTask {
    uint64_t byte_off; // of the unit's panel inside side 0's stream
    int32_t lrow, lcol, p0, k0;
    uint16_t h, w, ld, flags;
};
static_assert(sizeof(DenseTask) == 32, "DenseTask must be 32 bytes");

struct PackOptions {
    int sort_units      = 1;     // order the units of a block by the rows they act on (RUNS of the multi-RHS kernels); 0: leaf order
    bool generate_dense = false; // dense leaves with data0 == NULL are allowed: their panels are generated on the device
    int block_rows  = 0;     // 32, 64 or 128; 0 = automatic (128 for double, 64 for complex<double>)
    bool near_field = false; // build the multi-RHS near-field layout (NearFieldLayout; needs sort_units). Off by default: measured, DESIGN.md 6
    int nf_rows     = 0;     // rows of a near-field panel (rounded up to a multiple of 8; 0: one panel over all the rows of a block)
    int piece_cols  = 16;    // columns of a unit (<= 32); lowered automatically so that a unit fits a stage
    int stage_bytes = 24576; // bulk-copy granule of the coefficient stream, multiple of 16
    int cseg_bytes  = 4096;  // capacity of a stage's c segment, multiple of 16 (512 columns: stages of small-cluster runs still fill up)
    // Side 0 (target rows) may use smaller blocks than side 1 (0 = same as block_rows). Tried for the row strips of a
    // distributed operator (few target rows, all the source columns: 1024 APPLY blocks = 2.3 waves at 8 GPUs): halving
    // the target blocks speeds APPLY up by 8 % but the extra chunks cost the same in COMBINE (gpurun_out/t32), so the
    // default keeps one height; the knob stays for experiments.
    int target_block_rows = 0;
    // Tail split of side 0 (packer.cpp: make_blocks): quarter-height blocks for the rows of the last, partial round of
    // APPLY CTAs when the side has between 1 and 6 rounds of blocks. cta_slots = resident APPLY CTAs of the device.
    int tail_split = 1;
    int cta_slots  = 148 * 3;
};

// ---- multi-RHS near field (mkernels.cu, APPLY_M) ------------------------------------------------------------------------------
// The dense leaves of a target block (the near field: ~16 leaf clusters of 7 - 8 rows, each against ~27 source clusters that
// are mostly the SAME for the whole block) cost the multi-RHS APPLY pass far more than their 7 % of the coefficients: 8-row
// runs that straddle two row tiles, B rows (rows of the input matrix) fetched again by every leaf cluster. For the multi-RHS
// product 'N' they are therefore ALSO kept as ONE block-sparse panel per target block: all the rows of the block x the union
// of the columns its dense leaves touch (zeros where there is no leaf), in stages of the usual size — the full-height run path of
// APPLY_M, every B row fetched once per block. It is a second copy of the dense coefficients (x ~2.5 with the zeros), built
// on the device from the main stream the first time it is needed; the single-RHS kernels never see it.
struct NfTask { // one dense unit (or the part of it that falls into one near-field stage): h x w panel, main stream -> near-field stream
    uint64_t src_off, dst_off; // bytes
    uint16_t h, w, src_ld, dst_ld;
    uint32_t reserved[2];
};
static_assert(sizeof(NfTask) == 32, "NfTask must be 32 bytes");
struct NearFieldLayout {
    std::vector<BlockDesc> blocks; // only the target blocks that hold dense units
    std::vector<StageDesc> stages;
    std::vector<uint32_t> order;
    std::vector<unsigned char> aux_apply; // one run per stage: all the rows, K columns = rows of the input matrix
    std::vector<unsigned char> headers;   // 16 bytes per stage (StageHeader with no unit)
    std::vector<uint64_t> hdr_off;        // n_stages + 1 (= 16 st)
    std::vector<NfTask> tasks;
    uint64_t stream_bytes  = 0;
    uint32_t aux_max_bytes = 0;
    uint64_t coefficients  = 0; // panel entries, zeros included
    bool empty() const { return blocks.empty(); }
};

// std::vector that does not zero its elements on resize: the big side tables (hundreds of MB) are sized once and then filled by
// all the threads of the packer — the first touch of a page happens in the parallel copy, not in a serial memset.
template <typename T>
struct DefaultInitAllocator : std::allocator<T> {
    template <typename U>
    struct rebind {
        using other = DefaultInitAllocator<U>;
    };
    DefaultInitAllocator() = default;
    template <typename U>
    DefaultInitAllocator(const DefaultInitAllocator<U> &) {}
    template <typename U>
    void construct(U *ptr) noexcept(std::is_nothrow_default_constructible<U>::value) {
#ifdef HTB_POISON_RAW_VECTORS /* debug builds: an entry the packer forgets to write shows up in the digests of tools/pack_time.py */
        std::memset(static_cast<void *>(ptr), 0xAB, sizeof(U));
#else
        ::new (static_cast<void *>(ptr)) U;
#endif
    }
    template <typename U, typename... Args>
    void construct(U *ptr, Args &&...args) {
        ::new (static_cast<void *>(ptr)) U(std::forward<Args>(args)...);
    }
};
template <typename T>
using RawVector = std::vector<T, DefaultInitAllocator<T>>;

// Host description of one side, produced by the packer. Device copies are owned by the handle.
struct SideLayout {
    int n = 0; // length of the index space
    std::vector<BlockDesc> blocks;
    std::vector<StageDesc> stages;
    std::vector<uint32_t> order; // block ids, heaviest stream first
    // direction whose CONSUMER is this side (producers are the units of the other side)
    std::vector<CombineEntry> combine;
    std::vector<CombineDst> combine_dst;
    uint64_t stream_bytes = 0;
    uint64_t n_units      = 0;
    uint64_t part_base = 0, part_elems = 0; // PART[this side as consumer]
    uint64_t cs_base = 0, cs_elems = 0;     // CS[this side]
    bool any_twice = false;
    // multi-RHS tables (this side as CONSUMER for combine_m / partm)
    RawVector<MUnit> munits;
    std::vector<CombineEntry> combine_m; // src = PARTM offset, dst_first = TF offset, n_dst unused
    uint64_t partm_base = 0, partm_elems = 0; // in vectors, inside the multi-RHS scratch [TF | PARTM[0] | PARTM[1]]
    RawVector<unsigned char> aux_reduce, aux_apply; // per-stage aux records (runs + column tables), same offsets in both
    uint32_t aux_max_bytes = 0;                        // largest aux record of the side: sizes the aux part of a ring slot
    RawVector<DenseTask> dense_tasks;                 // side 0 only: dense units to generate on the device
    RawVector<DenseTask> lr_tasks;                    // units of low-rank leaves whose factors live in the device pool (lrow = leaf index, lcol = side)
};

} // namespace htb
#endif

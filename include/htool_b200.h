/*
 * htool_b200.h — C ABI of the B200-native H-matrix product.
 *
 * This is the drop-in boundary. The reference (htool-ddm/htool, header-only C++) has no FFI of its
 * own: its plugin API is the pair of abstract classes
 *     htool::VirtualLocalToLocalOperator<T>   include/htool/distributed_operator/interfaces/virtual_local_to_local_operator.hpp:8-35
 *     htool::VirtualGlobalToLocalOperator<T>  include/htool/distributed_operator/interfaces/virtual_global_to_local_operator.hpp:8-35
 * (three virtuals each: add_vector_product :16, add_matrix_product_row_major :25,
 * add_sub_matrix_product_to_local :33). The header-only shim in
 * htool_b200/cpp/htool_b200/ (operators.hpp) subclasses them and forwards every call to the entry points
 * below, so this file declares exactly what those virtuals (and the free functions
 * add_hmatrix_vector_product / add_hmatrix_matrix_product,
 * include/htool/hmatrix/linalg/add_hmatrix_vector_product.hpp:173,
 * include/htool/hmatrix/linalg/add_hmatrix_matrix_product.hpp:176) need.
 *
 * Conventions
 *   - plain pointers and sizes only; scalars are passed by pointer so one ABI serves double (8 B)
 *     and complex<double> (16 B, re/im interleaved like std::complex<double>);
 *   - every function returns an int status (HTB_OK == 0); htb_last_error() gives the message of the
 *     last failure on the calling thread. The reference itself never throws or returns codes — it logs
 *     through htool::Logger and goes on (add_hmatrix_vector_product.hpp:112-115) — so the shim logs
 *     the message through htool::Logger and returns;
 *   - vectors are in CLUSTER numbering, relative to the root block's target/source offsets, exactly
 *     like openmp_internal_add_hmatrix_vector_product (add_hmatrix_vector_product.hpp:107-170);
 *   - a handle owns one CUDA stream + workspaces: one product in flight per handle (same contract as
 *     every caller in the reference: one thread per MPI rank);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     HTB_ERR_CUDA.
 */
#ifndef HTOOL_B200_H
#define HTOOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTB_VERSION_MAJOR 0
#define HTB_VERSION_MINOR 1

/* status codes */
enum {
    HTB_OK              = 0,
    HTB_ERR_INVALID     = 1, /* bad argument (null pointer, negative size, unknown trans, ...) */
    HTB_ERR_UNSUPPORTED = 2, /* trans='T' with 'H' storage or trans='C' with 'S' storage: the reference logs
                                "Operation is not supported" (add_hmatrix_vector_product.hpp:112-115) */
    HTB_ERR_CUDA        = 3, /* CUDA runtime / driver failure, or no device */
    HTB_ERR_NCCL        = 4,
    HTB_ERR_ALLOC       = 5
};

/* coefficient type */
enum { HTB_DOUBLE = 0,
       HTB_COMPLEX_DOUBLE = 1 };

/* where the in/out vectors of a product live */
enum { HTB_MEM_HOST = 0,   /* pageable or pinned host memory: the library stages through pinned buffers */
       HTB_MEM_DEVICE = 1 /* device pointers on the handle's device: no copies */ };

/* htb_leaf.flags */
#define HTB_LEAF_APPLY_TRANSPOSED_TOO 0x1 /* leaf is in get_leaves_from(...).second (hmatrix.hpp:264-266) */
#define HTB_LEAF_DIAG_SYMMETRIC 0x2       /* dense leaf with get_symmetry()=='S' -> symv (add_hmatrix_vector_product.hpp:22-24) */
#define HTB_LEAF_DIAG_HERMITIAN 0x4       /* dense leaf with get_symmetry()=='H' -> hemv (:43-45) */
#define HTB_LEAF_UPLO_UPPER 0x8           /* get_UPLO()=='U' (else 'L') for the two flags above */

/* One leaf of the block tree (hmatrix.hpp:29-50). Offsets are in cluster numbering and RELATIVE to the
 * root block (target offset - root target offset), the same subtraction the reference does at
 * add_hmatrix_vector_product.hpp:148-150. */
typedef struct htb_leaf {
    int32_t row_offset;
    int32_t col_offset;
    int32_t nb_rows; /* m */
    int32_t nb_cols; /* n */
    int32_t rank;    /* -1: dense leaf; r >= 0: low-rank leaf (r == 0 contributes nothing, add_lrmat_vector_product.hpp:11) */
    int32_t flags;
    const void *data0; /* dense: A, m x n column-major, lda = m (matrix.hpp:20-26). low rank: U, m x r column-major (lrmat.hpp:19) */
    const void *data1; /* low rank: V, r x n column-major. dense: NULL */
} htb_leaf;

/* Everything htb_create needs; filled by htool_b200::flatten() from an htool::HMatrix. Host pointers
 * only need to stay valid during htb_create: all coefficients are copied to the device once. */
typedef struct htb_hmatrix_desc {
    int32_t dtype;    /* HTB_DOUBLE | HTB_COMPLEX_DOUBLE */
    int32_t nb_rows;  /* root target cluster size */
    int32_t nb_cols;  /* root source cluster size */
    int32_t row_offset; /* root target cluster offset (informational; used by the distributed entry points) */
    int32_t col_offset; /* root source cluster offset */
    char symmetry_for_leaves; /* root get_symmetry_for_leaves(): 'N' | 'S' | 'H' (hmatrix.hpp:217) */
    char uplo_for_leaves;     /* 'N' | 'L' | 'U' */
    char reserved[2];
    int32_t device; /* CUDA device ordinal; -1 = the calling thread's current device */
    int64_t nb_leaves;
    const htb_leaf *leaves;
} htb_hmatrix_desc;

typedef struct htb_operator *htb_handle;

/* Numbers describing the device-resident leaf store of a handle. */
typedef struct htb_info {
    int64_t nb_leaves, nb_dense_leaves, nb_low_rank_leaves, nb_leaves_applied_twice;
    int64_t coefficients;         /* C  = sum_dense m*n + sum_lowrank r*(m+n)   */
    int64_t coefficients_twice;   /* C_sym: coefficients of leaves applied twice */
    int64_t store_bytes;          /* device bytes of packed coefficients (== sizeof(T)*C + padding) */
    int64_t descriptor_bytes;     /* device bytes of work-item descriptors */
    int64_t workspace_bytes;      /* device bytes of scratch (t vectors, staging) */
    int32_t rank_min, rank_max;
    int32_t dtype, device;
    int32_t nb_rows, nb_cols;
    int32_t nb_target_blocks, nb_source_blocks;
    int32_t sm_count;
    int32_t dist_gather;          /* 0: no communicator, 1: NCCL gather of x, 2: peer-memory push over NVLink (htb_comm_init) */
} htb_info;

/* ---- life cycle ---------------------------------------------------------------------------------- */

/* Packs the leaves into the device-resident store (uploaded once per assembly). */
int htb_create(const htb_hmatrix_desc *desc, htb_handle *out);
int htb_destroy(htb_handle h);
int htb_get_info(htb_handle h, htb_info *info);

/* Launch on a caller-owned stream (a cudaStream_t passed as void*); NULL restores the handle's own
 * stream. Lets a caller time with its own events / order against its own work. */
int htb_set_stream(htb_handle h, void *cuda_stream);
/* Blocks until every product enqueued on the handle has finished. */
int htb_synchronize(htb_handle h);

/* Optional: page-locks a caller-owned host buffer (cudaHostRegister) so that the HTB_MEM_HOST entry points copy it
 * with the DMA engines directly instead of staging it through the handle's pinned buffers. The reference's callers
 * (HPDDM's Krylov vectors, wrapper_hpddm.hpp:118-124; DistributedOperator work buffers) reuse the same buffers for
 * every product, so one registration serves a whole solve. Buffers that are already page-locked (cudaMallocHost,
 * torch pin_memory) are detected without this call. Unregister before freeing the buffer. */
int htb_host_register(void *ptr, size_t bytes);
int htb_host_unregister(void *ptr);

/* ---- leaf assembly on the device (first step: the dense near-field leaves) ---------------------------- */

/* Built-in kernel functions, evaluated on the device with the reference's own formulas
 * (include/htool/testing/generator_test.hpp:155-205; Helmholtz exp(ikr)/(4 pi r) with a finite diagonal). */
enum { HTB_KERNEL_LAPLACE = 0,       /* 1 / (4 pi r)                          double          */
       HTB_KERNEL_LAPLACE_REG = 1,   /* 1 / (1e-5 + 4 pi r)                   double          */
       HTB_KERNEL_COMPLEX_REG = 2,   /* (1 + i) / (1e-5 + 4 pi r)             complex<double> */
       HTB_KERNEL_HERMITIAN_REG = 3, /* (1 + sign(x_t - x_s) i) / (1e-5 + 4 pi r)             */
       HTB_KERNEL_HELMHOLTZ = 4,     /* exp(i k r) / (4 pi r)                                 */
       HTB_KERNEL_COMPLEX = 5        /* (1 + i) / (4 pi r)                                    */ };
typedef struct htb_generator_desc {
    int32_t kernel;            /* HTB_KERNEL_* */
    int32_t spatial_dimension; /* 3 */
    double wavenumber;         /* k of HTB_KERNEL_HELMHOLTZ */
    const double *target_points; /* 3 doubles per row of the root block, in CLUSTER numbering (row i of the block <-> point of
                                    user index permutation[row_offset + i]); host memory, read during the call */
    const double *source_points; /* same for the columns */
} htb_generator_desc;
/* htb_create for an H-matrix whose dense leaves may come WITHOUT coefficients (htb_leaf.data0 == NULL, rank == -1): those
 * leaves are generated on the device, straight into the leaf store, from the built-in kernel function and the points —
 * what HMatrix::compute_dense_data (hmatrix.hpp:222-226) does per leaf on the host, in the batch shape of
 * VirtualDenseBlocksGenerator::copy_dense_blocks (virtual_dense_blocks_generator.hpp:12, tree_builder.hpp:650-665): the
 * list of blocks is the list of such leaves. Dense leaves that do carry data0 (e.g. admissible blocks whose compression
 * failed) and all low-rank leaves are packed from the host as usual. Symmetric / Hermitian diagonal leaves are rebuilt
 * from their UPLO triangle like host data. Real kernels reproduce compute_dense_data bit for bit. */
int htb_create_generated(const htb_hmatrix_desc *desc, const htb_generator_desc *generator, htb_handle *out);
/* ---- leaf assembly on the device (second step: the admissible blocks) -------------------------------- */

/* rank of a leaf that is an ADMISSIBLE block still to be compressed (data0 == data1 == NULL): htb_create_compressed runs the
 * reference's adaptive cross approximation on the device for all such leaves at once. */
#define HTB_RANK_COMPRESS (-2)
typedef struct htb_compression_info {
    int64_t nb_blocks;          /* admissible blocks handed to the device                                             */
    int64_t nb_failed;          /* of which stored as dense leaves ("false positives", tree_builder.hpp:619-625)      */
    int64_t coefficients;       /* sum r (m + n) over the compressed blocks                                           */
    int64_t pool_bytes;         /* device bytes of the factor pool while the store was built                          */
    int32_t rank_min, rank_max; /* over the compressed blocks                                                         */
    double seconds_aca;         /* the ACA kernels                                                                    */
    double seconds_total;       /* the whole call (ACA + packing + upload + generation of the dense leaves + copies)  */
    double seconds_aca_team[3]; /* the ACA kernels by team size: 512, 128, 32 threads per block                       */
    int64_t nb_blocks_team[3];  /* admissible blocks of each team size                                                */
    double seconds_layout;      /* host: layout of the store (ranks known)                                            */
    double seconds_upload;      /* descriptors + streams to the device (the panels travel as zeros)                   */
    double seconds_fill;        /* device: generation of the dense leaves + copy of the factors into the streams      */
    double seconds_prepare;     /* host: list of the blocks, device allocations, points to the device                 */
    double seconds_compress;    /* pool allocation + the ACA kernels + ranks back to the host                         */
    double seconds_store;       /* layout + upload + fill + scratch allocations (the htb_create part)                 */
} htb_compression_info;
/* htb_create for an H-matrix whose block cluster tree is known but whose leaves hold NO coefficients: dense leaves
 * (rank == -1, data0 == NULL) are generated as in htb_create_generated, admissible leaves (rank == HTB_RANK_COMPRESS) are
 * compressed on the device by sympartialACA (include/htool/hmatrix/lrmat/sympartialACA.hpp:41-216, the reference's default
 * compressor, tree_builder.hpp:385) with the reference's stopping criterion at `epsilon` — what
 * HMatrixTreeBuilder::openmp_compute_blocks (tree_builder.hpp:604-666) does on the host: HMatrix::compute_low_rank_data
 * (hmatrix.hpp:228-237) per admissible block, a dense leaf instead when the compression fails (:619-625). Pivots, ranks and
 * factors are those of the reference (bit for bit for the real kernel functions when the host BLAS computes axpy with a
 * rounded product and a rounded sum, as the OpenBLAS of this image does; option aca_fma_axpy = 1 for FMA builds). Leaves
 * that carry data are packed from the host as usual. All the built-in kernel functions: the complex ones follow the reference's
 * std::complex<double> instantiation (std::abs, the compiler's complex division and product, zaxpy as y += ar x, y += ai (i x));
 * those without transcendental calls reproduce the reference's factors bit for bit, Helmholtz to the last ulps of sincos. */
int htb_create_compressed(const htb_hmatrix_desc *desc, const htb_generator_desc *generator, double epsilon, htb_handle *out);
/* Ranks the device found, in the order of the descriptor's leaves: the rank for compressed leaves, -1 for dense leaves
 * (including admissible blocks whose compression failed), the descriptor's rank for leaves that carried data. */
int htb_get_leaf_ranks(htb_handle h, int32_t *ranks, int64_t nb_leaves);
int htb_get_compression_info(htb_handle h, htb_compression_info *info);

/* Test / debug: copies the first `bytes` bytes of one side's packed stream (layout: htool_b200/csrc/store.hpp) back from
 * the device, e.g. to compare device-generated leaves with htb_pack_host of the host-generated ones. */
int htb_download_store(htb_handle h, int side, void *dst, int64_t bytes);

/* ---- products ------------------------------------------------------------------------------------- */

/* out <- beta*out + alpha*op(H)*in, op = N | T | C. Replaces openmp_internal_add_hmatrix_vector_product
 * (add_hmatrix_vector_product.hpp:107-170) as called from LocalToLocalHMatrix::add_vector_product
 * (local_to_local_operators/hmatrix.hpp:27-29) and RestrictedGlobalToLocalHMatrix::local_add_vector_product
 * (global_to_local_operators/hmatrix.hpp:27-29).
 * trans == 'N': in has nb_cols entries, out nb_rows; otherwise swapped.
 * HTB_MEM_HOST: returns after out is complete. HTB_MEM_DEVICE: asynchronous on the handle's stream. */
int htb_add_vector_product(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mem_kind);

/* out <- beta*out + alpha*op(H)*in for mu right-hand sides stored ROW-MAJOR (mu contiguous). Replaces
 * openmp_internal_add_hmatrix_matrix_product_row_major(trans,'N',...)
 * (add_hmatrix_matrix_product_row_major.hpp:112-178). */
int htb_add_matrix_product_row_major(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, int mem_kind);

/* User-numbering front ends, add_hmatrix_vector_product (add_hmatrix_vector_product.hpp:173-197) and
 * add_hmatrix_matrix_product with column-major B and C (add_hmatrix_matrix_product.hpp:176-205): the
 * permutation gathers/scatters of cluster_node.hpp:150-175 run on the device.
 * target_permutation / source_permutation: the cluster permutations restricted to the root block,
 * already shifted so that entries are in [0, nb_rows) / [0, nb_cols) (perm[offset+i]-offset). */
int htb_set_permutations(htb_handle h, const int32_t *target_permutation, const int32_t *source_permutation);
int htb_add_vector_product_user_numbering(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mem_kind);
int htb_add_matrix_product_user_numbering(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, int mem_kind);

/* ---- distributed (one process per GPU, NCCL) ------------------------------------------------------- */

#define HTB_NCCL_UNIQUE_ID_BYTES 128
/* Rank 0 calls this and broadcasts the 128 bytes to the other ranks by any means (MPI_Bcast, torch.distributed). */
int htb_nccl_get_unique_id(void *id128);
/* Attaches an NCCL communicator to the handle. partition_offsets has world_size+1 entries: rank r owns
 * rows/cols [partition_offsets[r], partition_offsets[r+1]) of the global cluster numbering
 * (PartitionFromCluster::get_offset_of_partition, partition_from_cluster.hpp:22-23). The handle's
 * H-matrix must be that rank's row strip (built with target_partition_number = rank,
 * distributed_operator/utility.hpp:56). */
int htb_comm_init(htb_handle h, const void *id128, int world_size, int rank, const int32_t *partition_offsets);
int htb_comm_destroy(htb_handle h);
/* Local-to-local distributed product, mu >= 1 row-major. Replaces
 * internal_add_distributed_operator_vector_product_local_to_local
 * (add_distributed_operator_vector_product_local_to_local.hpp:19-89) and its row-major matrix twin
 * (add_distributed_operator_matrix_product_row_major_local_to_local.hpp:25-95).
 *   trans == 'N': out_local <- beta*out_local + alpha*H_strip*allgather(in_local). The gather (MPI_Allgatherv at
 *     linalg/utility.hpp:27) runs on a second stream and overlaps with the leaves whose source range lies inside
 *     the rank's own partition.
 *   trans == 'T' | 'C': z = alpha*op(H_strip)^T*in_local has the global length; slice r of z is sent to rank r
 *     (MPI_Alltoallv at :77, grouped ncclSend/ncclRecv here) and out_local <- beta*out_local + sum_r z_r[own slice],
 *     added in rank order like the reference's axpy loop (:83-86). */
int htb_dist_add_product_local_to_local(htb_handle h, char trans, const void *alpha, const void *in_local, const void *beta, void *out_local, int mu, int mem_kind);
/* Global-to-global distributed product in partition (cluster) numbering, square operator, mu >= 1 row-major. Replaces
 * internal_add_distributed_operator_vector_product_global_to_global
 * (add_distributed_operator_vector_product_global_to_global.hpp:18-85) and the row-major matrix twin
 * (add_distributed_operator_matrix_product_row_major_global_to_global.hpp:18-84).
 *   'N': own rows of out <- beta*out + alpha*H_strip*in_global, then the gather of out (MPI_Allgatherv, :75).
 *   'T' | 'C': partial = alpha*op(H_strip)^T*in_global[own rows]; out <- allreduce_sum(partial) + beta*out (:77-83). */
int htb_dist_add_product_global_to_global(htb_handle h, char trans, const void *alpha, const void *in_global, const void *beta, void *out_global, int mu, int mem_kind);

/* ---- device-resident Krylov loop ------------------------------------------------------------------------ */

/* Restarted GMRES on A x = rhs with A = the handle's operator, all vectors resident in HBM. Stands for
 * DDM::solve with "-hpddm_schwarz_method none" (solvers/ddm.hpp:134-193), i.e. HPDDM's unpreconditioned GMRES calling
 * HPDDMOperator::GMV (wrappers/wrapper_hpddm.hpp:102-145) per iteration; HPDDM is not vendored in the reference, so the
 * Krylov arithmetic restates the published algorithm with HPDDM's defaults (restart 40, 100 iterations, tolerance 1e-6
 * relative to ||rhs||, classical Gram-Schmidt). With a communicator (htb_comm_init) rhs / x are the rank's LOCAL slices,
 * the product is htb_dist_add_product_local_to_local and the inner products are summed over the ranks; without one the
 * operator must be square. x holds the initial guess on entry and the solution on return. Collective when distributed. */
enum { HTB_GMRES_CGS = 0, HTB_GMRES_CGS2 = 1 };
typedef struct htb_gmres_options {
    int32_t restart;               /* Krylov vectors kept before restarting */
    int32_t max_iterations;        /* total inner iterations (products, not counting the residual of each cycle) */
    double tolerance;              /* stop when ||rhs - A x|| <= tolerance * ||rhs|| (recurrence estimate) */
    int32_t orthogonalization;     /* HTB_GMRES_CGS (HPDDM default) | HTB_GMRES_CGS2 */
    int32_t verbosity;             /* 0 silent, 1 summary, 2 every iteration (stderr) */
    int32_t compute_true_residual; /* one more product at the end */
    int32_t reserved;
} htb_gmres_options;
typedef struct htb_gmres_result {
    int32_t iterations, converged, matvecs, reserved;
    double relative_residual;      /* from the Givens recurrence */
    double true_relative_residual; /* ||rhs - A x|| / ||rhs||, -1 if not computed */
} htb_gmres_result;
int htb_gmres_default_options(htb_gmres_options *options);
int htb_gmres(htb_handle h, const void *rhs, void *x, const htb_gmres_options *options, htb_gmres_result *result, int mem_kind);

/* ---- misc ------------------------------------------------------------------------------------------ */

const char *htb_last_error(void);
int htb_device_count(int *count);
/* Number of kernel launches issued by the handle since creation (bench.py reports it). */
int htb_launch_count(htb_handle h, int64_t *count);
/* Per-kernel timing with CUDA events recorded on the launching stream around every launch (off by
 * default; serialises nothing that was not already serial). htb_get_pass_times synchronises, returns the
 * accumulated device milliseconds and launch counts per kernel kind since the last call, and resets them. */
enum { HTB_PASS_REDUCE = 0, /* t = V x / op(U)^T x / op(A)^T x, streams one side of the store */
       HTB_PASS_COMBINE = 1, /* folds per-chunk partials and replicates them into the c-stream slots */
       HTB_PASS_APPLY = 2,   /* y = beta y + alpha (U t + A x) / op(V)^T t, streams one side of the store */
       HTB_PASS_OTHER = 3,   /* permutations, scaling */
       HTB_PASS_KINDS = 4 };
int htb_profile_passes(htb_handle h, int enable);
int htb_get_pass_times(htb_handle h, double ms[HTB_PASS_KINDS], int64_t launches[HTB_PASS_KINDS]);
/* Tunables for experiments; unknown keys return HTB_ERR_INVALID. Read when the store is packed (set them before htb_create):
 * block_rows, piece_cols, stage_bytes, cseg_bytes, ring_stages, reduce_ring_stages, evict_first, upload_chunk_mb, sort_units,
 * upload_headers_only (device assembly: only the stage headers travel), m_near_field / m_nf_rows (second, block-sparse copy of the
 * dense leaves for the multi-RHS product 'N'; off: measured, no gain). Read by htb_create_compressed: aca_fma_axpy (the BLAS the
 * reference was linked with fuses its axpy), aca_dots (0: guarded warp-level dot products, 1: in order, 2: replay always),
 * aca_rank_guess (terms per block the factor pool is first sized for). Read at product time: mrhs_min (right-hand sides from which
 * the tensor-core path is used; 0 = automatic), m_fast_tall, m_b_global, m_small_runs, m_reduce_split, m_stage_input (multi-RHS
 * kernel variants), zero_copy, fused_symmetric, pdl, dist_p2p. The table is PROCESS-WIDE (one per loaded library, not per handle): a
 * value read at product time applies to every live handle; pdl is latched by the most recent htb_create. */
int htb_set_option(const char *key, int64_t value);
int htb_get_option(const char *key, int64_t *value);

/* ---- packer introspection (host only, no CUDA) ----------------------------------------------------- */

/* The bytes and tables htb_create would upload for one side of the store (0: target rows = U panels + dense
 * leaves, 1: source columns = V^T panels), layouts in htool_b200/csrc/store.hpp. The combine tables are those of
 * the direction whose CONSUMER is this side. It exists so the CPU test-suite can check the stream format without
 * a GPU; it computes no product. */
typedef struct htb_packed_side {
    int32_t n;                /* length of the side's index space */
    int32_t n_blocks;
    int64_t n_stages;
    int64_t n_combine;
    int64_t n_combine_dst;
    int64_t stream_bytes;
    int64_t scratch_elems;    /* elements of ONE scratch copy: PART[0] | PART[1] | CS[0] | CS[1] */
    int64_t cs_base, cs_elems;     /* c-stream of this side inside a scratch copy */
    int64_t part_base, part_elems; /* partials of the direction whose consumer is this side */
    int32_t piece_cols;       /* effective unit width */
    int32_t block_rows;       /* effective block height */
    int64_t n_munits, n_combine_m; /* multi-RHS side tables (store.hpp: MUnit, CombineEntry) */
    int64_t mscratch_elems;   /* vectors of ONE multi-RHS scratch copy: TF | PARTM[0] | PARTM[1] */
    const void *blocks;       /* n_blocks x 32 B  (BlockDesc) */
    const void *stages;       /* n_stages x 32 B  (StageDesc) */
    const void *order;        /* n_blocks x uint32: launch order, heaviest block first */
    const void *combine;      /* n_combine x 16 B (CombineEntry) */
    const void *combine_dst;  /* n_combine_dst x 8 B (CombineDst) */
    const void *stream;       /* stream_bytes */
    const void *munits;       /* n_munits x 16 B (MUnit), stage order */
    const void *combine_m;    /* n_combine_m x 16 B (CombineEntry: src = PARTM offset, dst_first = TF offset) */
    void *owner;              /* opaque, released by htb_pack_free */
    int64_t aux_bytes;        /* multi-RHS aux records of the side's stages (store.hpp: AuxHeader | RunDesc[] | column table) */
    const void *aux_reduce;   /* column entry = scratch vector receiving the column's REDUCE_M result */
    const void *aux_apply;    /* column entry = scratch vector (or input row | bit 31) the column multiplies in APPLY_M */
    int64_t n_dense_tasks;    /* side 0, option pack_generate_dense = 1: dense units of leaves without data0 (store.hpp: DenseTask, 32 B) */
    const void *dense_tasks;
    int64_t n_lowrank_tasks;  /* either side, same option: units of low-rank leaves without data0 / data1 (their factors are in the
                               * device pool of htb_create_compressed); DenseTask with lrow = leaf index, lcol = side, p0 = first row of
                               * the panel inside the leaf's factor, k0 = first term */
    const void *lowrank_tasks;
    int64_t header_bytes;     /* when NO leaf carries data (device assembly): the stage headers packed back to back, what travels instead of */
    const void *headers;      /* the stream (header of stage st = headers[header_offsets[st] .. header_offsets[st + 1]) -> stream + stage byte_off); */
    const void *header_offsets; /* (n_stages + 1) x uint64; NULL / 0 when some leaf carries host data                                      */
} htb_packed_side;
int htb_pack_host(const htb_hmatrix_desc *desc, int side, htb_packed_side *out);
int htb_pack_free(htb_packed_side *packed);

#ifdef __cplusplus
}
#endif
#endif /* HTOOL_B200_H */

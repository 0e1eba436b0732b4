// htool_b200/csrc/handle.hpp — the object behind htb_handle (include/htool_b200.h).
#ifndef HTB_HANDLE_HPP
#define HTB_HANDLE_HPP

#include "kernels.cuh"
#include "mkernels.cuh"
#include "packer.hpp"
#include <cuda_runtime.h>
#include <htool_b200.h>
#include <memory>
#include <string>
#include <vector>

namespace htb {
struct DistState; // dist.cu
// Distributed 'N' product: the source-side REDUCE is launched in two parts, the blocks that lie inside the
// rank's own partition (they only need the local x) and the others, which wait for the allgather.
struct DistSplit {
    const uint32_t *order_local  = nullptr;
    const uint32_t *order_remote = nullptr;
    int n_local = 0, n_remote = 0;
    cudaEvent_t gather_done = nullptr; // NCCL gather: the remote part waits for this event
    // peer-memory gather (preferred): ONE launch over order_all (local blocks first); remote blocks wait in the kernel
    // for the arrival flag of the rank that owns their slice of x
    const uint32_t *order_all            = nullptr;
    int n_all                            = 0;
    const uint32_t *owner                = nullptr;
    const unsigned long long *flags      = nullptr;
    unsigned long long epoch             = 0;
    int world                            = 1;
};
}

struct htb_operator {
    int device = 0, dtype = 0, sm_count = 0;
    size_t esize = 8;
    int nb_rows = 0, nb_cols = 0, row_offset = 0, col_offset = 0;
    char symmetry = 'N', uplo = 'N';
    htb::SideDevice side[2];
    std::vector<htb::BlockDesc> host_blocks[2]; // host copies, used to split passes by index range (dist.cu)
    std::vector<uint32_t> host_order[2];
    htb::LaunchConfig launch_cfg;
    std::vector<void *> owned; // device allocations of the store
    void *d_scratch      = nullptr;
    uint64_t scratch_elems = 0;
    // multi-RHS scratch ([TF | PARTM[0] | PARTM[1]] x vector stride), allocated at the first multi-RHS product
    std::vector<int32_t> leaf_ranks;      // htb_create_compressed: ranks per leaf of the descriptor
    htb_compression_info compression{};
    double create_seconds[3] = {0., 0., 0.}; // layout (host), upload, device fill (generation + factor copies)
    // multi-RHS near field (store.hpp, NearFieldLayout): host layout kept until the first multi-RHS product 'N' builds the device copy
    std::unique_ptr<htb::NearFieldLayout> nf_host;
    htb::SideDevice nf{};
    bool nf_ready = false, nf_failed = false;
    uint64_t nf_bytes = 0;
    void *d_mscratch        = nullptr;
    void *d_mstage          = nullptr; // multi-RHS: the current column group of the input, rows padded to the B-ring stride
    size_t mstage_cap       = 0;
    uint64_t mscratch_elems = 0; // vectors per copy
    int mscratch_vs         = 0; // vector stride it was allocated for
    bool needs_second_copy  = false;
    bool fused_symmetric    = true;  // symmetric storage: second application fused into the first APPLY pass
    bool m_path_ok          = false; // double, and the multi-RHS kernels fit the shared memory with the current options
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // host-pointer entry points: pinned + device staging, grown on demand
    void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    std::vector<cudaEvent_t> chunk_events; // D2H pipeline of the host-pointer entry points
    // user-numbering front ends
    void *d_perm[2]  = {nullptr, nullptr};
    void *d_work_in = nullptr, *d_work_out = nullptr;
    size_t work_cap = 0;
    int64_t launches = 0;
    // optional per-kernel timing (htb_profile_passes)
    bool profiling = false;
    struct TimedLaunch {
        cudaEvent_t start, stop;
        int kind;
    };
    std::vector<TimedLaunch> timed;
    size_t store_bytes = 0, descriptor_bytes = 0, workspace_bytes = 0;
    uint64_t side_stream_bytes[2] = {0, 0};
    int64_t generated_dense_units = 0; // dense units generated on the device (htb_create_generated)
    htb_info info{};
    htb::DistState *dist = nullptr;
    // Krylov workspace (gmres.cu), grown on demand and kept between solves
    void *d_krylov = nullptr, *h_krylov = nullptr;
    size_t krylov_cap = 0, krylov_pin_cap = 0;
};

namespace htb {
extern thread_local std::string g_last_error;
int fail(int status, const std::string &msg);
int64_t option_value(const char *key); // htb_set_option / htb_get_option table
int cuda_fail(cudaError_t e, const char *what);
int product_device(htb_operator *h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, const DistSplit *split = nullptr);
int ensure_staging(htb_operator *h, size_t in_bytes, size_t out_bytes);
// pageable host <-> device through the pinned staging buffers, in 1 MiB chunks so that the CPU copy of one chunk
// overlaps the DMA of the previous one
int staged_h2d(htb_operator *h, void *dev, void *pinned, const void *host, size_t bytes, cudaStream_t st);
int staged_d2h(htb_operator *h, void *host, void *pinned, const void *dev, size_t bytes, cudaStream_t st); // returns when host is complete
void *mapped_device_pointer(const void *host, size_t bytes); // device address of a page-locked mapped host buffer covering [host, host + bytes), or nullptr
void dist_destroy(htb_operator *h);
int dist_gather_mode(const htb_operator *h); // 0 none, 1 NCCL, 2 peer memory
} // namespace htb
#endif

// htool_b200/csrc/dist.cu — one process per GPU: NCCL allgather of x overlapped with the local leaves.
//
// Replaces, for trans == 'N', internal_add_distributed_operator_vector_product_local_to_local
// (include/htool/distributed_operator/linalg/add_distributed_operator_vector_product_local_to_local.hpp:19-59)
// and its row-major matrix twin (add_distributed_operator_matrix_product_row_major_local_to_local.hpp:25-66):
// there, local_to_global() is one MPI_Allgatherv of x (linalg/utility.hpp:11-28) and the product only
// starts when it has returned. Here the gather runs on its own stream while the source blocks that lie
// inside the rank's own partition are already being reduced (t = V x needs nothing remote for them).
// Partition sizes differ between ranks (RegularSplitting gives the remainder to the last child,
// clustering/implementations/partitioning.hpp:241-246). Rows are owned: no reduction is needed for 'N'.
//
// The gather of x, two implementations:
//   * peer memory (default on one box): every rank maps the x buffers and arrival flags of its peers (CUDA IPC over
//     NVLink / NVSwitch). A PUSH kernel stores the rank's slice of x straight into every peer's buffer and then
//     releases an epoch flag on each peer; the REDUCE kernel of the product is ONE launch whose own-partition
//     blocks start at once and whose other blocks wait, block by block, for the flag of the rank that owns their
//     slice — transfer and compute overlap at block granularity and no collective call sits on the critical path.
//     Buffers are double-buffered by epoch parity and the push waits for its peers' previous epoch (WAR guard).
//   * NCCL (fallback, and the T / C and global-to-global paths): a group of in-place ncclBroadcasts, one per rank,
//     on a second stream; the remote part of the REDUCE waits for an event.
#include "handle.hpp"

#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h> // types only: the library is bound at run time, see NcclApi

namespace htb {

// NCCL is bound lazily with dlopen instead of being a link-time dependency: a process that also imports
// PyTorch already holds PyTorch's own libnccl.so.2, and two different NCCL builds under one SONAME cannot
// coexist. An already loaded libnccl.so.2 is reused (RTLD_NOLOAD); otherwise the system library is opened.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *)                                                                   = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                            = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                                       = nullptr;
    ncclResult_t (*GroupStart)()                                                                                  = nullptr;
    ncclResult_t (*GroupEnd)()                                                                                    = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)        = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                     = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                           = nullptr;
    const char *(*GetErrorString)(ncclResult_t)                                                                   = nullptr;
    bool ok = false;
    std::string error;
};

static NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib)
            lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) {
            api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(lib, name);
            if (!p)
                api.error = std::string("libnccl.so.2 lacks ") + name;
            return p;
        };
        api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GroupStart     = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd       = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.Broadcast      = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.AllReduce      = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.Send           = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv           = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok             = api.error.empty();
    });
    return api;
}

struct DistState {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    std::vector<int32_t> offsets;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t in_ready = nullptr, gather_done = nullptr;
    void *d_xglobal = nullptr; // gathered x ('N'), full-length partial result (T / C)
    size_t xglobal_cap = 0;
    void *d_rbuf = nullptr; // T / C local-to-local: the world slices received for the own partition
    size_t rbuf_cap = 0;
    void *d_old = nullptr; // T / C global-to-global with beta != 0: out before the product
    size_t old_cap = 0;
    uint32_t *d_order_local = nullptr, *d_order_remote = nullptr;
    int n_local = 0, n_remote = 0;
    // ---- peer-memory gather ------------------------------------------------------------------------------------------
    bool p2p = false;                  // decided collectively at htb_comm_init
    std::string p2p_why;               // why not, when disabled
    void *xg[2] = {nullptr, nullptr};  // own gathered-x buffers (epoch parity)
    size_t xg_cap = 0;
    unsigned long long *flags = nullptr; // own arrival flags [world]
    unsigned int *arrive      = nullptr; // push kernel: CTAs done
    std::vector<void *> peer_xg[2];      // peers' buffers mapped here (own entry = own pointer)
    std::vector<unsigned long long *> peer_flags;
    void **d_peer_xg[2]                 = {nullptr, nullptr}; // the same pointer tables on the device
    unsigned long long **d_peer_flags   = nullptr;
    unsigned long long epoch            = 0;
    uint32_t *d_order_all = nullptr, *d_owner = nullptr;
    void *d_hbuf = nullptr; // handle exchange
    // the WAR guard of the peer-memory gather assumes that a rank's push is ordered after its previous product; that is
    // stream order as long as the caller keeps one stream, and this event when htb_set_stream switched streams in between
    cudaEvent_t product_done    = nullptr;
    cudaStream_t product_stream = nullptr;
    // ---- small all-reduce over peer memory (inner products of the Krylov loop, gmres.cu) ------------------------------------
    // every rank owns an inbox [2 parities][world][kRedMax doubles] and arrival flags [world]; a reduction = every rank
    // stores its values into slot [parity][rank] of EVERY inbox, releases flag[rank] = epoch on every peer, waits for the
    // world flags of its own inbox and sums the slots in rank order (same order on every rank: identical results)
    double *red_inbox = nullptr;
    unsigned long long *red_flags = nullptr;
    std::vector<double *> peer_red_inbox;
    std::vector<unsigned long long *> peer_red_flags;
    double **d_peer_red_inbox             = nullptr;
    unsigned long long **d_peer_red_flags = nullptr;
    unsigned long long red_epoch          = 0;
    bool red_p2p                          = false;
};
constexpr int kRedMax = 256; // doubles per reduction (restart 40, complex, two Gram-Schmidt passes: 2 * 2 * 42 = 168)

static int nccl_fail(ncclResult_t r, const char *what) {
    return fail(HTB_ERR_NCCL, std::string(what) + ": " + nccl().GetErrorString(r));
}

static void close_peer_buffers(DistState *d) {
    for (int k = 0; k < 2; k++) {
        for (size_t r = 0; r < d->peer_xg[k].size(); r++)
            if (static_cast<int>(r) != d->rank && d->peer_xg[k][r])
                cudaIpcCloseMemHandle(d->peer_xg[k][r]);
        d->peer_xg[k].clear();
    }
}

int dist_world(const htb_operator *h) { return h->dist ? h->dist->world : 0; }

int dist_allreduce_small(htb_operator *h, double *dev, int n, cudaStream_t st);
// in-place sum over the ranks of n doubles on the device (inner products of the Krylov loop, gmres.cu)
int dist_allreduce_sum(htb_operator *h, double *dev, size_t n, cudaStream_t st) {
    DistState *d = h->dist;
    if (!d || d->world <= 1 || n == 0)
        return HTB_OK;
    if (d->red_p2p && n <= static_cast<size_t>(kRedMax)) {
        // (defined below) no collective call, no NCCL kernel: one small CTA per rank over the peer mappings
        return dist_allreduce_small(h, dev, static_cast<int>(n), st);
    }
    ncclResult_t r = nccl().AllReduce(dev, dev, n, ncclDouble, ncclSum, d->comm, st);
    if (r != ncclSuccess)
        return fail(HTB_ERR_NCCL, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
    return HTB_OK;
}

int dist_gather_mode(const htb_operator *h) { return !h->dist ? 0 : (h->dist->p2p ? 2 : 1); }

void dist_destroy(htb_operator *h) {
    DistState *d = h->dist;
    if (!d)
        return;
    if (d->comm_stream)
        cudaStreamSynchronize(d->comm_stream);
    // (not h->stream: a caller-owned stream set with htb_set_stream may be gone by now; the event covers the last product)
    if (d->product_done) {
        cudaEventSynchronize(d->product_done);
        cudaEventDestroy(d->product_done);
    }
    if (h->own_stream)
        cudaStreamSynchronize(h->own_stream);
    close_peer_buffers(d);
    for (size_t r = 0; r < d->peer_flags.size(); r++)
        if (static_cast<int>(r) != d->rank && d->peer_flags[r])
            cudaIpcCloseMemHandle(d->peer_flags[r]);
    d->peer_flags.clear();
    for (size_t r = 0; r < d->peer_red_inbox.size(); r++)
        if (static_cast<int>(r) != d->rank && d->peer_red_inbox[r])
            cudaIpcCloseMemHandle(d->peer_red_inbox[r]);
    for (size_t r = 0; r < d->peer_red_flags.size(); r++)
        if (static_cast<int>(r) != d->rank && d->peer_red_flags[r])
            cudaIpcCloseMemHandle(d->peer_red_flags[r]);
    for (void *p : {static_cast<void *>(d->red_inbox), static_cast<void *>(d->red_flags), static_cast<void *>(d->d_peer_red_inbox), static_cast<void *>(d->d_peer_red_flags)})
        if (p)
            cudaFree(p);
    for (void *p : {d->xg[0], d->xg[1], static_cast<void *>(d->flags), static_cast<void *>(d->arrive), static_cast<void *>(d->d_peer_xg[0]), static_cast<void *>(d->d_peer_xg[1]),
                    static_cast<void *>(d->d_peer_flags), static_cast<void *>(d->d_order_all), static_cast<void *>(d->d_owner), d->d_hbuf})
        if (p)
            cudaFree(p);
    if (d->comm)
        nccl().CommDestroy(d->comm);
    for (void *p : {static_cast<void *>(d->d_xglobal), static_cast<void *>(d->d_rbuf), static_cast<void *>(d->d_old), static_cast<void *>(d->d_order_local), static_cast<void *>(d->d_order_remote)})
        if (p)
            cudaFree(p);
    if (d->in_ready)
        cudaEventDestroy(d->in_ready);
    if (d->gather_done)
        cudaEventDestroy(d->gather_done);
    if (d->comm_stream)
        cudaStreamDestroy(d->comm_stream);
    delete d;
    h->dist = nullptr;
}

// out[i] = beta out[i] + sum_r rbuf[r][i], r = 0 .. world-1 in rank order: the scal + world axpys that follow the
// MPI_Alltoallv of the reference (add_distributed_operator_vector_product_local_to_local.hpp:79-86), same order.
// One ELEMENT per thread (a complex element = both of its words: the beta product reads re and im of the old value, so
// the two words must be read before either is written — out may be updated in place).
template <bool CPLX>
__global__ void sum_slices_kernel(double *out, const double *rbuf, long long n_elems, int world, double beta_re, double beta_im, int beta_is_zero) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_elems)
        return;
    constexpr int W       = CPLX ? 2 : 1;
    const long long i     = e * W;
    const long long n_dbl = n_elems * W;
    double v[W];
#pragma unroll
    for (int j = 0; j < W; j++)
        v[j] = 0.;
    if (!beta_is_zero) {
        if (CPLX) {
            const double re = out[i], im = out[i + 1];
            v[0]     = beta_re * re - beta_im * im;
            v[W - 1] = beta_re * im + beta_im * re;
        } else
            v[0] = beta_re * out[i];
    }
    for (int r = 0; r < world; r++)
#pragma unroll
        for (int j = 0; j < W; j++)
            v[j] += rbuf[static_cast<size_t>(r) * n_dbl + i + j];
#pragma unroll
    for (int j = 0; j < W; j++)
        out[i + j] = v[j];
}

// out[i] = sum[i] + beta old[i] (global-to-global T / C: the axpy after MPI_Allreduce,
// add_distributed_operator_vector_product_global_to_global.hpp:77-83). One element per thread: old may alias out.
template <bool CPLX>
__global__ void add_scaled_kernel(double *out, const double *sum, const double *old, long long n_elems, double beta_re, double beta_im) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_elems)
        return;
    if (CPLX) {
        const long long i = 2 * e;
        const double re = old[i], im = old[i + 1];
        const double s0 = sum[i], s1 = sum[i + 1];
        out[i]     = s0 + (beta_re * re - beta_im * im);
        out[i + 1] = s1 + (beta_re * im + beta_im * re);
    } else
        out[e] = sum[e] + beta_re * old[e];
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// PUSH: this rank's slice of x (n8 8-byte words) is stored into the gathered-x buffer of EVERY rank of the box (own
// included) at the slice's global offset, over NVLink peer mappings; when all CTAs are done the last one releases
// flags[rank] = epoch on every peer. Replaces the MPI_Allgatherv of linalg/utility.hpp:27 on the 'N' path.
// WAR guard: buffer (epoch & 1) was last read by the peers' products of epoch - 2; a peer that has published epoch - 1
// has finished that product (its push is stream-ordered after it), so wait for flags >= epoch - 1 first.
__global__ void push_x_kernel(const unsigned long long *src, size_t n8, void *const *peer_xg, size_t my_off8, unsigned long long *const *peer_flags, int world, int rank,
                              unsigned long long epoch, const unsigned long long *own_flags, unsigned int *arrive) {
    if (epoch >= 2 && static_cast<int>(threadIdx.x) < world && static_cast<int>(threadIdx.x) != rank)
        while (ld_acquire_sys_u64(own_flags + threadIdx.x) < epoch - 1)
            __nanosleep(64);
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const unsigned long long v = src[i];
        for (int p = 0; p < world; p++)
            static_cast<unsigned long long *>(peer_xg[p])[my_off8 + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(arrive, 1u);
        if (prev == gridDim.x - 1) {
            *arrive = 0; // next launch is stream-ordered after this one
            __threadfence_system();
            for (int p = 0; p < world; p++)
                st_release_sys_u64(peer_flags[p] + rank, epoch);
        }
    }
}

// In-place sum over the ranks of n <= kRedMax doubles, one CTA. WAR safety of the two inbox parities: a rank stores epoch e
// into parity e & 1 only after it has seen every peer's flag e - 1 (inside its reduction e - 1), and a peer publishes e - 1
// only after its reduction e - 2 — the last reader of that parity — has finished (stream order).
__global__ void allreduce_small_kernel(double *vals, int n, double *const *peer_inbox, unsigned long long *const *peer_flags, const double *own_inbox, const unsigned long long *own_flags, int world,
                                       int rank, unsigned long long epoch) {
    const int t        = threadIdx.x;
    const size_t slot  = (static_cast<size_t>(epoch & 1ull) * world + rank) * kRedMax;
    if (t < n) {
        const double v = vals[t];
        for (int p = 0; p < world; p++)
            peer_inbox[p][slot + t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < world)
        st_release_sys_u64(peer_flags[t] + rank, epoch);
    if (t < world)
        while (ld_acquire_sys_u64(own_flags + t) < epoch)
            __nanosleep(32);
    __syncthreads();
    if (t < n) {
        double s = 0.;
        for (int r = 0; r < world; r++)
            s += own_inbox[(static_cast<size_t>(epoch & 1ull) * world + r) * kRedMax + t];
        vals[t] = s;
    }
}

static cudaError_t grow_device(void **p, size_t *cap, size_t need) {
    if (need <= *cap)
        return cudaSuccess;
    if (*p)
        cudaFree(*p);
    *p   = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess)
        *cap = need;
    return e;
}

int dist_allreduce_small(htb_operator *h, double *dev, int n, cudaStream_t st) {
    DistState *d = h->dist;
    const unsigned long long epoch = ++d->red_epoch;
    allreduce_small_kernel<<<1, kRedMax, 0, st>>>(dev, n, d->d_peer_red_inbox, d->d_peer_red_flags, d->red_inbox, d->red_flags, d->world, d->rank, epoch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return cuda_fail(e, "allreduce_small_kernel");
    h->launches++;
    return HTB_OK;
}

} // namespace htb

using namespace htb;

#define HTB_CUDA(call)                    \
    do {                                  \
        cudaError_t e__ = (call);         \
        if (e__ != cudaSuccess)           \
            return cuda_fail(e__, #call); \
    } while (0)
#define HTB_NCCL(call)                    \
    do {                                  \
        ncclResult_t r__ = (call);        \
        if (r__ != ncclSuccess)           \
            return nccl_fail(r__, #call); \
    } while (0)

// ---- peer-memory plumbing (collective calls: every rank of the communicator makes them in the same order) -------------
// min over the ranks of a local 0/1 verdict: the ranks must AGREE on the gather implementation
static int agree(DistState *d, bool local_ok, bool *all_ok) {
    double v = local_ok ? 1. : 0.;
    double *dv = reinterpret_cast<double *>(static_cast<char *>(d->d_hbuf) + size_t(d->world) * sizeof(cudaIpcMemHandle_t));
    HTB_CUDA(cudaMemcpyAsync(dv, &v, sizeof(double), cudaMemcpyHostToDevice, d->comm_stream));
    if (d->world > 1)
        HTB_NCCL(nccl().AllReduce(dv, dv, 1, ncclDouble, ncclMin, d->comm, d->comm_stream));
    HTB_CUDA(cudaMemcpyAsync(&v, dv, sizeof(double), cudaMemcpyDeviceToHost, d->comm_stream));
    HTB_CUDA(cudaStreamSynchronize(d->comm_stream));
    *all_ok = v > 0.5;
    return HTB_OK;
}

// Every rank exports `own` (a cudaMalloc allocation) and maps the allocations of the others: mapped[r] on return
// (mapped[rank] = own). *ok is the local verdict; mappings that were opened stay in `mapped` for the caller to close.
static int exchange_and_map(DistState *d, void *own, std::vector<void *> &mapped, bool *ok) {
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    std::vector<cudaIpcMemHandle_t> handles(d->world);
    *ok = true;
    std::memset(handles.data(), 0, hs * d->world);
    if (cudaIpcGetMemHandle(&handles[d->rank], own) != cudaSuccess) {
        cudaGetLastError();
        *ok = false;
    }
    char *hb = static_cast<char *>(d->d_hbuf);
    HTB_CUDA(cudaMemcpyAsync(hb + hs * d->rank, &handles[d->rank], hs, cudaMemcpyHostToDevice, d->comm_stream));
    if (d->world > 1) {
        HTB_NCCL(nccl().GroupStart());
        for (int r = 0; r < d->world; r++)
            HTB_NCCL(nccl().Broadcast(hb + hs * r, hb + hs * r, hs, ncclChar, r, d->comm, d->comm_stream));
        HTB_NCCL(nccl().GroupEnd());
    }
    HTB_CUDA(cudaMemcpyAsync(handles.data(), hb, hs * d->world, cudaMemcpyDeviceToHost, d->comm_stream));
    HTB_CUDA(cudaStreamSynchronize(d->comm_stream));
    mapped.assign(d->world, nullptr);
    mapped[d->rank] = own;
    for (int r = 0; r < d->world && *ok; r++) {
        if (r == d->rank)
            continue;
        if (cudaIpcOpenMemHandle(&mapped[r], handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            d->p2p_why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError());
            mapped[r]  = nullptr;
            *ok        = false;
        }
    }
    return HTB_OK;
}

static int upload_table(void **dev_table, const void *host_table, size_t bytes) {
    if (!*dev_table)
        HTB_CUDA(cudaMalloc(dev_table, bytes));
    HTB_CUDA(cudaMemcpy(*dev_table, host_table, bytes, cudaMemcpyHostToDevice));
    return HTB_OK;
}

// flags + tables, once per communicator (htb_comm_init)
static int p2p_init(htb_operator *h, DistState *d) {
    int rc;
    bool ok = option_value("dist_p2p") != 0 && d->world <= 32, all = false;
    if (!ok)
        d->p2p_why = d->world > 32 ? "more than 32 ranks" : "disabled (option dist_p2p = 0)";
    HTB_CUDA(cudaMalloc(&d->d_hbuf, size_t(d->world) * sizeof(cudaIpcMemHandle_t) + 16));
    HTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d->flags), sizeof(unsigned long long) * d->world));
    HTB_CUDA(cudaMemset(d->flags, 0, sizeof(unsigned long long) * d->world));
    HTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d->arrive), sizeof(unsigned int)));
    HTB_CUDA(cudaMemset(d->arrive, 0, sizeof(unsigned int)));
    if ((rc = agree(d, ok, &all)) != HTB_OK)
        return rc;
    if (!all)
        return HTB_OK; // NCCL gather
    std::vector<void *> mapped;
    if ((rc = exchange_and_map(d, d->flags, mapped, &ok)) != HTB_OK)
        return rc;
    d->peer_flags.resize(d->world);
    for (int r = 0; r < d->world; r++)
        d->peer_flags[r] = static_cast<unsigned long long *>(mapped[r]);
    if (ok && (rc = upload_table(reinterpret_cast<void **>(&d->d_peer_flags), d->peer_flags.data(), sizeof(void *) * d->world)) != HTB_OK)
        return rc;
    if ((rc = agree(d, ok, &all)) != HTB_OK)
        return rc;
    d->p2p = all;
    if (d->p2p) {
        // inbox + flags of the small all-reduce, mapped on every rank like the gather flags (collective sequence: every rank
        // goes through the same exchanges whatever its local verdict)
        bool rok = d->world <= kRedMax;
        if (rok && cudaMalloc(reinterpret_cast<void **>(&d->red_inbox), sizeof(double) * 2 * d->world * kRedMax) != cudaSuccess) {
            cudaGetLastError();
            rok = false;
        }
        if (rok && cudaMalloc(reinterpret_cast<void **>(&d->red_flags), sizeof(unsigned long long) * d->world) != cudaSuccess) {
            cudaGetLastError();
            rok = false;
        }
        if (rok) {
            HTB_CUDA(cudaMemset(d->red_inbox, 0, sizeof(double) * 2 * d->world * kRedMax));
            HTB_CUDA(cudaMemset(d->red_flags, 0, sizeof(unsigned long long) * d->world));
        }
        std::vector<void *> m1, m2;
        bool ok1 = rok, ok2 = rok;
        if ((rc = exchange_and_map(d, rok ? static_cast<void *>(d->red_inbox) : static_cast<void *>(d->flags), m1, &ok1)) != HTB_OK)
            return rc;
        if ((rc = exchange_and_map(d, rok ? static_cast<void *>(d->red_flags) : static_cast<void *>(d->flags), m2, &ok2)) != HTB_OK)
            return rc;
        rok = rok && ok1 && ok2;
        d->peer_red_inbox.resize(d->world);
        d->peer_red_flags.resize(d->world);
        for (int r = 0; r < d->world; r++) {
            d->peer_red_inbox[r] = static_cast<double *>(m1[r]);
            d->peer_red_flags[r] = static_cast<unsigned long long *>(m2[r]);
        }
        if (!rok) { // entries that point at the stand-in allocation must not be closed as if they were ours
            d->peer_red_inbox[d->rank] = nullptr;
            d->peer_red_flags[d->rank] = nullptr;
        }
        if (rok && (rc = upload_table(reinterpret_cast<void **>(&d->d_peer_red_inbox), d->peer_red_inbox.data(), sizeof(void *) * d->world)) != HTB_OK)
            return rc;
        if (rok && (rc = upload_table(reinterpret_cast<void **>(&d->d_peer_red_flags), d->peer_red_flags.data(), sizeof(void *) * d->world)) != HTB_OK)
            return rc;
        if ((rc = agree(d, rok, &all)) != HTB_OK)
            return rc;
        d->red_p2p = all;
    }
    // launch order of the single REDUCE: own-partition blocks first; owners of every block's index range
    const std::vector<BlockDesc> &blocks = h->host_blocks[1];
    std::vector<uint32_t> owner(blocks.size(), 0xFFFFFFFFu), order_all;
    const int lo = d->offsets[d->rank], hi = d->offsets[d->rank + 1];
    auto owner_of = [&](int idx) {
        int q = static_cast<int>(std::upper_bound(d->offsets.begin(), d->offsets.end(), idx) - d->offsets.begin()) - 1;
        return std::min(std::max(q, 0), d->world - 1);
    };
    for (size_t b = 0; b < blocks.size(); b++) {
        const BlockDesc &bd = blocks[b];
        if (bd.nrows > 0 && !(bd.row_start >= lo && bd.row_start + bd.nrows <= hi))
            owner[b] = static_cast<uint32_t>(owner_of(bd.row_start)) | (static_cast<uint32_t>(owner_of(bd.row_start + bd.nrows - 1)) << 16);
    }
    for (uint32_t b : h->host_order[1])
        if (owner[b] == 0xFFFFFFFFu)
            order_all.push_back(b);
    for (uint32_t b : h->host_order[1])
        if (owner[b] != 0xFFFFFFFFu)
            order_all.push_back(b);
    HTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d->d_owner), std::max<size_t>(1, owner.size()) * 4));
    HTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d->d_order_all), std::max<size_t>(1, order_all.size()) * 4));
    HTB_CUDA(cudaMemcpy(d->d_owner, owner.data(), owner.size() * 4, cudaMemcpyHostToDevice));
    HTB_CUDA(cudaMemcpy(d->d_order_all, order_all.data(), order_all.size() * 4, cudaMemcpyHostToDevice));
    return HTB_OK;
}

// (re)allocation of the double-buffered gathered x when a product needs more room: collective, rare (first product,
// or a larger mu). Falls back to the NCCL gather for good if any rank cannot map its peers.
static int p2p_ensure_buffers(htb_operator *h, DistState *d, size_t bytes) {
    if (!d->p2p || bytes <= d->xg_cap)
        return HTB_OK;
    int rc;
    bool ok = true, all = false;
    HTB_CUDA(cudaStreamSynchronize(h->stream));
    close_peer_buffers(d);
    if ((rc = agree(d, true, &all)) != HTB_OK) // barrier: nobody still maps the buffers about to be freed
        return rc;
    for (int k = 0; k < 2; k++) {
        if (d->xg[k])
            cudaFree(d->xg[k]);
        d->xg[k] = nullptr;
    }
    d->xg_cap = 0;
    for (int k = 0; k < 2 && ok; k++) {
        if (cudaMalloc(&d->xg[k], bytes) != cudaSuccess) {
            cudaGetLastError();
            d->p2p_why = "cudaMalloc of the gathered-x buffers";
            ok         = false;
        }
    }
    for (int k = 0; k < 2; k++) {
        bool okk = ok;
        void *own = ok ? d->xg[k] : static_cast<void *>(d->flags); // keep the collective sequence identical on every rank
        if ((rc = exchange_and_map(d, own, d->peer_xg[k], &okk)) != HTB_OK)
            return rc;
        ok = ok && okk;
        if (ok && (rc = upload_table(reinterpret_cast<void **>(&d->d_peer_xg[k]), d->peer_xg[k].data(), sizeof(void *) * d->world)) != HTB_OK)
            return rc;
    }
    if ((rc = agree(d, ok, &all)) != HTB_OK)
        return rc;
    if (all)
        d->xg_cap = bytes;
    else {
        close_peer_buffers(d);
        d->p2p = false;
    }
    return HTB_OK;
}

extern "C" {

int htb_nccl_get_unique_id(void *id128) {
    if (!id128)
        return fail(HTB_ERR_INVALID, "null argument");
    static_assert(sizeof(ncclUniqueId) == HTB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    if (!nccl().ok)
        return fail(HTB_ERR_NCCL, nccl().error);
    ncclUniqueId id;
    HTB_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return HTB_OK;
}

int htb_comm_init(htb_handle h, const void *id128, int world_size, int rank, const int32_t *partition_offsets) {
    if (!h || !id128 || !partition_offsets || world_size < 1 || rank < 0 || rank >= world_size)
        return fail(HTB_ERR_INVALID, "invalid argument");
    for (int r = 0; r < world_size; r++)
        if (partition_offsets[r] > partition_offsets[r + 1])
            return fail(HTB_ERR_INVALID, "partition offsets must be non-decreasing");
    if (partition_offsets[0] != 0 || partition_offsets[world_size] != h->nb_cols)
        return fail(HTB_ERR_INVALID, "partition offsets must cover [0, nb_cols) of the row strip");
    if (partition_offsets[rank + 1] - partition_offsets[rank] != h->nb_rows || partition_offsets[rank] != h->row_offset - h->col_offset)
        return fail(HTB_ERR_INVALID, "the handle is not the row strip of this rank's partition");
    if (!nccl().ok)
        return fail(HTB_ERR_NCCL, nccl().error);
    int prev = 0;
    cudaGetDevice(&prev);
    HTB_CUDA(cudaSetDevice(h->device));
    dist_destroy(h);
    auto *d    = new DistState();
    h->dist    = d;
    d->world   = world_size;
    d->rank    = rank;
    d->offsets.assign(partition_offsets, partition_offsets + world_size + 1);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    HTB_NCCL(nccl().CommInitRank(&d->comm, world_size, id, rank));
    HTB_CUDA(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
    HTB_CUDA(cudaEventCreateWithFlags(&d->in_ready, cudaEventDisableTiming));
    HTB_CUDA(cudaEventCreateWithFlags(&d->gather_done, cudaEventDisableTiming));
    HTB_CUDA(cudaEventCreateWithFlags(&d->product_done, cudaEventDisableTiming));
    // source blocks entirely inside the own partition vs the rest, each keeping the heaviest-first order
    std::vector<uint32_t> local, remote;
    const int lo = d->offsets[rank], hi = d->offsets[rank + 1];
    for (uint32_t b : h->host_order[1]) {
        const BlockDesc &bd = h->host_blocks[1][b];
        if (bd.row_start >= lo && bd.row_start + bd.nrows <= hi)
            local.push_back(b);
        else
            remote.push_back(b);
    }
    d->n_local  = static_cast<int>(local.size());
    d->n_remote = static_cast<int>(remote.size());
    HTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d->d_order_local), std::max<size_t>(1, local.size()) * 4));
    HTB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d->d_order_remote), std::max<size_t>(1, remote.size()) * 4));
    HTB_CUDA(cudaMemcpy(d->d_order_local, local.data(), local.size() * 4, cudaMemcpyHostToDevice));
    HTB_CUDA(cudaMemcpy(d->d_order_remote, remote.data(), remote.size() * 4, cudaMemcpyHostToDevice));
    const int prc = p2p_init(h, d);
    cudaSetDevice(prev);
    return prc;
}

int htb_comm_destroy(htb_handle h) {
    if (!h)
        return fail(HTB_ERR_INVALID, "null handle");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    dist_destroy(h);
    cudaSetDevice(prev);
    return HTB_OK;
}

// in-place gather of the per-rank segments of a global vector: a group of broadcasts, one per rank (partition sizes differ)
static int gather_segments(DistState *d, char *base, size_t es, cudaStream_t st) {
    if (d->world <= 1)
        return HTB_OK;
    HTB_NCCL(nccl().GroupStart());
    for (int r = 0; r < d->world; r++) {
        const size_t count = size_t(d->offsets[r + 1] - d->offsets[r]) * es;
        if (count == 0)
            continue;
        char *seg = base + size_t(d->offsets[r]) * es;
        HTB_NCCL(nccl().Broadcast(seg, seg, count, ncclChar, r, d->comm, st));
    }
    HTB_NCCL(nccl().GroupEnd());
    return HTB_OK;
}

namespace {
struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        cudaSetDevice(dev);
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};
bool scalar_is_zero(const htb_operator *h, const void *s) {
    const double *b = static_cast<const double *>(s);
    return b[0] == 0. && (h->dtype == HTB_DOUBLE || b[1] == 0.);
}
} // namespace

int htb_dist_add_product_local_to_local(htb_handle h, char trans, const void *alpha, const void *in_local, const void *beta, void *out_local, int mu, int mem_kind) {
    if (!h || !alpha || !beta || !in_local || !out_local || mu < 1)
        return fail(HTB_ERR_INVALID, "invalid argument");
    DistState *d = h->dist;
    if (!d)
        return fail(HTB_ERR_INVALID, "htb_comm_init has not been called");
    DeviceGuard guard(h->device);

    const size_t es = h->esize * mu;
    const size_t n_local = h->nb_rows, n_global = h->nb_cols;
    cudaStream_t st = h->stream;
    char *xg        = nullptr;
    void *dout      = out_local;
    const bool beta_zero = scalar_is_zero(h, beta);
    int rc;
    if (trans == 'N') {
        if ((rc = p2p_ensure_buffers(h, d, n_global * es)) != HTB_OK)
            return rc;
        // zero copy (single RHS, page-locked mapped host vectors, peer-memory gather): the push kernel reads x_local from
        // the host buffer and the APPLY epilogue writes y_local into it, no staging copies
        void *zc_in = nullptr, *zc_out = nullptr;
        if (mem_kind == HTB_MEM_HOST && mu == 1 && d->p2p && option_value("zero_copy") != 0) {
            zc_in  = mapped_device_pointer(in_local, n_local * es);
            zc_out = mapped_device_pointer(out_local, n_local * es);
            if (!zc_in || !zc_out)
                zc_in = zc_out = nullptr;
        }
        if (zc_in) {
            mem_kind = HTB_MEM_DEVICE;
            in_local = zc_in;
            dout     = zc_out;
        } else if (mem_kind == HTB_MEM_HOST) {
            if ((rc = ensure_staging(h, n_local * es, n_local * es)) != HTB_OK)
                return rc;
            if (!beta_zero && (rc = staged_h2d(h, h->d_out, h->h_out, out_local, n_local * es, st)) != HTB_OK)
                return rc;
            dout = h->d_out;
        }
        const bool zero_copy = zc_in != nullptr;
        DistSplit split;
        split.world = d->world;
        if (d->p2p) {
            // peer-memory gather: push the own slice into every rank's buffer of this epoch's parity, then ONE reduce launch
            const unsigned long long epoch = ++d->epoch;
            const int k = static_cast<int>(epoch & 1ull);
            if (d->product_stream && d->product_stream != st) // the caller switched streams: order this push after the previous product
                HTB_CUDA(cudaStreamWaitEvent(st, d->product_done, 0));
            xg          = static_cast<char *>(d->xg[k]);
            const void *src = in_local;
            if (mem_kind == HTB_MEM_HOST) {
                if ((rc = staged_h2d(h, xg + size_t(d->offsets[d->rank]) * es, h->h_in, in_local, n_local * es, st)) != HTB_OK)
                    return rc;
                src = xg + size_t(d->offsets[d->rank]) * es;
            }
            const size_t n8 = n_local * es / 8;
            const int grid  = static_cast<int>(std::min<size_t>(std::max<size_t>(1, (n8 + 1023) / 1024), 2 * size_t(h->sm_count)));
            push_x_kernel<<<grid, 256, 0, st>>>(static_cast<const unsigned long long *>(src), n8, d->d_peer_xg[k], size_t(d->offsets[d->rank]) * es / 8, d->d_peer_flags, d->world, d->rank, epoch,
                                                d->flags, d->arrive);
            HTB_CUDA(cudaGetLastError());
            h->launches++;
            split.order_all = d->d_order_all;
            split.n_all     = d->n_local + d->n_remote;
            split.owner     = d->d_owner;
            split.flags     = d->flags;
            split.epoch     = epoch;
        } else {
            HTB_CUDA(grow_device(&d->d_xglobal, &d->xglobal_cap, n_global * es));
            xg = static_cast<char *>(d->d_xglobal);
            if (mem_kind == HTB_MEM_HOST) {
                if ((rc = staged_h2d(h, xg + size_t(d->offsets[d->rank]) * es, h->h_in, in_local, n_local * es, st)) != HTB_OK)
                    return rc;
            } else {
                HTB_CUDA(cudaMemcpyAsync(xg + size_t(d->offsets[d->rank]) * es, in_local, n_local * es, cudaMemcpyDeviceToDevice, st));
            }
            // allgather of x on the communication stream
            HTB_CUDA(cudaEventRecord(d->in_ready, st));
            HTB_CUDA(cudaStreamWaitEvent(d->comm_stream, d->in_ready, 0));
            if ((rc = gather_segments(d, xg, es, d->comm_stream)) != HTB_OK)
                return rc;
            HTB_CUDA(cudaEventRecord(d->gather_done, d->comm_stream));
            split.order_local  = d->d_order_local;
            split.order_remote = d->d_order_remote;
            split.n_local      = d->n_local;
            split.n_remote     = d->n_remote;
            split.gather_done  = d->gather_done;
        }
        if ((rc = product_device(h, 'N', alpha, xg, beta, dout, mu, &split)) != HTB_OK)
            return rc;
        if (d->p2p) {
            HTB_CUDA(cudaEventRecord(d->product_done, st));
            d->product_stream = st;
        }
        if (zero_copy) { // the caller's host vector is the output: complete before returning
            HTB_CUDA(cudaStreamSynchronize(st));
            return HTB_OK;
        }
    } else {
        // T / C (add_distributed_operator_vector_product_local_to_local.hpp:47-87): z = alpha op(H_strip)^T x_local has the
        // GLOBAL length; slice r of z goes to rank r (MPI_Alltoallv there, grouped ncclSend / ncclRecv here), and the owner
        // adds the world slices it received in rank order to beta out_local.
        const void *din = in_local;
        if (mem_kind == HTB_MEM_HOST) {
            if ((rc = ensure_staging(h, n_local * es, n_local * es)) != HTB_OK)
                return rc;
            if ((rc = staged_h2d(h, h->d_in, h->h_in, in_local, n_local * es, st)) != HTB_OK)
                return rc;
            if (!beta_zero && (rc = staged_h2d(h, h->d_out, h->h_out, out_local, n_local * es, st)) != HTB_OK)
                return rc;
            din  = h->d_in;
            dout = h->d_out;
        }
        HTB_CUDA(grow_device(&d->d_rbuf, &d->rbuf_cap, size_t(d->world) * n_local * es));
        HTB_CUDA(grow_device(&d->d_xglobal, &d->xglobal_cap, n_global * es));
        xg = static_cast<char *>(d->d_xglobal);
        alignas(16) const double zero[2] = {0., 0.};
        if ((rc = product_device(h, trans, alpha, din, zero, xg, mu)) != HTB_OK)
            return rc;
        char *rbuf = static_cast<char *>(d->d_rbuf);
        HTB_NCCL(nccl().GroupStart());
        for (int r = 0; r < d->world; r++) {
            const size_t scount = size_t(d->offsets[r + 1] - d->offsets[r]) * es;
            if (scount)
                HTB_NCCL(nccl().Send(xg + size_t(d->offsets[r]) * es, scount, ncclChar, r, d->comm, st));
            if (n_local)
                HTB_NCCL(nccl().Recv(rbuf + size_t(r) * n_local * es, n_local * es, ncclChar, r, d->comm, st));
        }
        HTB_NCCL(nccl().GroupEnd());
        const long long nd = static_cast<long long>(n_local * es / h->esize); // elements (a complex element is one work item)
        if (nd) {
            const double *b   = static_cast<const double *>(beta);
            const unsigned grid = static_cast<unsigned>((nd + 255) / 256);
            if (h->dtype == HTB_DOUBLE)
                sum_slices_kernel<false><<<grid, 256, 0, st>>>(static_cast<double *>(dout), static_cast<const double *>(d->d_rbuf), nd, d->world, b[0], 0., beta_zero ? 1 : 0);
            else
                sum_slices_kernel<true><<<grid, 256, 0, st>>>(static_cast<double *>(dout), static_cast<const double *>(d->d_rbuf), nd, d->world, b[0], b[1], beta_zero ? 1 : 0);
            HTB_CUDA(cudaGetLastError());
            h->launches++;
        }
    }
    if (mem_kind == HTB_MEM_HOST)
        return staged_d2h(h, out_local, h->h_out, h->d_out, n_local * es, st);
    return HTB_OK;
}

int htb_dist_add_product_global_to_global(htb_handle h, char trans, const void *alpha, const void *in_global, const void *beta, void *out_global, int mu, int mem_kind) {
    if (!h || !alpha || !beta || !in_global || !out_global || mu < 1)
        return fail(HTB_ERR_INVALID, "invalid argument");
    DistState *d = h->dist;
    if (!d)
        return fail(HTB_ERR_INVALID, "htb_comm_init has not been called");
    DeviceGuard guard(h->device);

    const size_t es = h->esize * mu;
    const size_t n_local = h->nb_rows, n_global = h->nb_cols; // square operator partitioned the same way on both sides
    const size_t my_off  = size_t(d->offsets[d->rank]) * es;
    cudaStream_t st = h->stream;
    const bool beta_zero = scalar_is_zero(h, beta);
    const bool host      = mem_kind == HTB_MEM_HOST;
    int rc;
    if (host && (rc = ensure_staging(h, n_global * es, n_global * es)) != HTB_OK)
        return rc;
    char *dout = host ? static_cast<char *>(h->d_out) : static_cast<char *>(out_global);
    if (trans == 'N') {
        // y_local = beta out[own rows] + alpha H_strip x, then the gather of y (MPI_Allgatherv,
        // add_distributed_operator_vector_product_global_to_global.hpp:43-50,74-76)
        const void *din = in_global;
        if (host) {
            if ((rc = staged_h2d(h, h->d_in, h->h_in, in_global, n_global * es, st)) != HTB_OK)
                return rc;
            if (!beta_zero && (rc = staged_h2d(h, dout + my_off, static_cast<char *>(h->h_out) + my_off, static_cast<const char *>(out_global) + my_off, n_local * es, st)) != HTB_OK)
                return rc;
            din = h->d_in;
        }
        if ((rc = product_device(h, 'N', alpha, din, beta, dout + my_off, mu)) != HTB_OK)
            return rc;
        if ((rc = gather_segments(d, dout, es, st)) != HTB_OK)
            return rc;
    } else {
        // partial = alpha op(H_strip)^T in[own rows] (global length), MPI_Allreduce(SUM), out = sum + beta out_before (:51-57,77-83)
        const char *din = static_cast<const char *>(in_global) + my_off;
        if (host) {
            if ((rc = staged_h2d(h, h->d_in, h->h_in, din, n_local * es, st)) != HTB_OK)
                return rc;
            din = static_cast<const char *>(h->d_in);
        }
        HTB_CUDA(grow_device(&d->d_xglobal, &d->xglobal_cap, n_global * es));
        alignas(16) const double zero[2] = {0., 0.};
        if ((rc = product_device(h, trans, alpha, din, zero, d->d_xglobal, mu)) != HTB_OK)
            return rc;
        const size_t nd = n_global * es / sizeof(double);
        if (beta_zero) {
            if (d->world > 1)
                HTB_NCCL(nccl().AllReduce(d->d_xglobal, dout, nd, ncclDouble, ncclSum, d->comm, st));
            else
                HTB_CUDA(cudaMemcpyAsync(dout, d->d_xglobal, n_global * es, cudaMemcpyDeviceToDevice, st));
        } else {
            if (d->world > 1)
                HTB_NCCL(nccl().AllReduce(d->d_xglobal, d->d_xglobal, nd, ncclDouble, ncclSum, d->comm, st));
            const void *old = out_global;
            if (host) {
                HTB_CUDA(grow_device(&d->d_old, &d->old_cap, n_global * es));
                if ((rc = staged_h2d(h, d->d_old, h->h_out, out_global, n_global * es, st)) != HTB_OK)
                    return rc;
                old = d->d_old;
            }
            const double *b     = static_cast<const double *>(beta);
            const long long ne  = static_cast<long long>(n_global * es / h->esize); // elements
            const unsigned grid = static_cast<unsigned>((ne + 255) / 256);
            if (h->dtype == HTB_DOUBLE)
                add_scaled_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<double *>(dout), static_cast<const double *>(d->d_xglobal), static_cast<const double *>(old), ne, b[0], 0.);
            else
                add_scaled_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<double *>(dout), static_cast<const double *>(d->d_xglobal), static_cast<const double *>(old), ne, b[0], b[1]);
            HTB_CUDA(cudaGetLastError());
            h->launches++;
        }
    }
    if (host)
        return staged_d2h(h, out_global, h->h_out, dout, n_global * es, st);
    return HTB_OK;
}

} // extern "C"

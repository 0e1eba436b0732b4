// htool_b200/csrc/dist.cu — one process per GPU: NCCL allgather of x overlapped with the local leaves.
//
// Replaces, for trans == 'N', internal_add_distributed_operator_vector_product_local_to_local
// (include/htool/distributed_operator/linalg/add_distributed_operator_vector_product_local_to_local.hpp:19-59)
// and its row-major matrix twin (add_distributed_operator_matrix_product_row_major_local_to_local.hpp:25-66):
// there, local_to_global() is one MPI_Allgatherv of x (linalg/utility.hpp:11-28) and the product only
// starts when it has returned. Here the gather runs on its own stream while the source blocks that lie
// inside the rank's own partition are already being reduced (t = V x needs nothing remote for them).
// Partition sizes differ between ranks (RegularSplitting gives the remainder to the last child,
// clustering/implementations/partitioning.hpp:241-246), so the gather is a group of in-place broadcasts,
// one per rank, rather than an equal-count ncclAllGather. Rows are owned: no reduction is needed.
#include "handle.hpp"

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
    void *d_xglobal = nullptr;
    size_t xglobal_cap = 0;
    uint32_t *d_order_local = nullptr, *d_order_remote = nullptr;
    int n_local = 0, n_remote = 0;
};

static int nccl_fail(ncclResult_t r, const char *what) {
    return fail(HTB_ERR_NCCL, std::string(what) + ": " + nccl().GetErrorString(r));
}

void dist_destroy(htb_operator *h) {
    DistState *d = h->dist;
    if (!d)
        return;
    if (d->comm_stream)
        cudaStreamSynchronize(d->comm_stream);
    if (d->comm)
        nccl().CommDestroy(d->comm);
    for (void *p : {static_cast<void *>(d->d_xglobal), static_cast<void *>(d->d_order_local), static_cast<void *>(d->d_order_remote)})
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
    cudaSetDevice(prev);
    return HTB_OK;
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

int htb_dist_add_product_local_to_local(htb_handle h, const void *alpha, const void *in_local, const void *beta, void *out_local, int mu, int mem_kind) {
    if (!h || !alpha || !beta || !in_local || !out_local || mu < 1)
        return fail(HTB_ERR_INVALID, "invalid argument");
    DistState *d = h->dist;
    if (!d)
        return fail(HTB_ERR_INVALID, "htb_comm_init has not been called");
    int prev = 0;
    cudaGetDevice(&prev);
    HTB_CUDA(cudaSetDevice(h->device));
    struct Restore {
        int dev;
        ~Restore() { cudaSetDevice(dev); }
    } restore{prev};

    const size_t es = h->esize * mu;
    const size_t n_local = h->nb_rows, n_global = h->nb_cols;
    if (n_global * es > d->xglobal_cap) {
        if (d->d_xglobal)
            cudaFree(d->d_xglobal);
        d->d_xglobal   = nullptr;
        d->xglobal_cap = 0;
        HTB_CUDA(cudaMalloc(&d->d_xglobal, n_global * es));
        d->xglobal_cap = n_global * es;
    }
    cudaStream_t st = h->stream;
    char *xg        = static_cast<char *>(d->d_xglobal);
    void *dout      = out_local;
    const double *b = static_cast<const double *>(beta);
    const bool beta_zero = b[0] == 0. && (h->dtype == HTB_DOUBLE || b[1] == 0.);
    if (mem_kind == HTB_MEM_HOST) {
        int rc = ensure_staging(h, n_local * es, n_local * es);
        if (rc != HTB_OK)
            return rc;
        if ((rc = staged_h2d(h, xg + size_t(d->offsets[d->rank]) * es, h->h_in, in_local, n_local * es, st)) != HTB_OK)
            return rc;
        if (!beta_zero && (rc = staged_h2d(h, h->d_out, h->h_out, out_local, n_local * es, st)) != HTB_OK)
            return rc;
        dout = h->d_out;
    } else {
        HTB_CUDA(cudaMemcpyAsync(xg + size_t(d->offsets[d->rank]) * es, in_local, n_local * es, cudaMemcpyDeviceToDevice, st));
    }
    // allgather of x on the communication stream
    HTB_CUDA(cudaEventRecord(d->in_ready, st));
    HTB_CUDA(cudaStreamWaitEvent(d->comm_stream, d->in_ready, 0));
    if (d->world > 1) {
        HTB_NCCL(nccl().GroupStart());
        for (int r = 0; r < d->world; r++) {
            const size_t count = size_t(d->offsets[r + 1] - d->offsets[r]) * es;
            if (count == 0)
                continue;
            char *seg = xg + size_t(d->offsets[r]) * es;
            HTB_NCCL(nccl().Broadcast(seg, seg, count, ncclChar, r, d->comm, d->comm_stream));
        }
        HTB_NCCL(nccl().GroupEnd());
    }
    HTB_CUDA(cudaEventRecord(d->gather_done, d->comm_stream));

    DistSplit split;
    split.order_local  = d->d_order_local;
    split.order_remote = d->d_order_remote;
    split.n_local      = d->n_local;
    split.n_remote     = d->n_remote;
    split.gather_done  = d->gather_done;
    int rc = product_device(h, 'N', alpha, xg, beta, dout, mu, &split);
    if (rc != HTB_OK)
        return rc;
    if (mem_kind == HTB_MEM_HOST)
        return staged_d2h(h, out_local, h->h_out, h->d_out, n_local * es, st);
    return HTB_OK;
}

} // extern "C"

// htool_b200/csrc/capi.cu — implementation of the C ABI declared in include/htool_b200.h.
//
// A handle owns: the device-resident leaf store (two stream-ordered slabs, store.hpp), the scratch for
// the t / z vectors, a CUDA stream, pinned staging buffers for the host-pointer entry points and,
// optionally, an NCCL communicator (dist.cu). There is no CPU path: every compute entry point needs a
// CUDA device and fails with HTB_ERR_CUDA otherwise.
#include "aca.cuh"
#include "generate.cuh"
#include "handle.hpp"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>

namespace htb {

thread_local std::string g_last_error;

int fail(int status, const std::string &msg) {
    g_last_error = msg;
    return status;
}
int cuda_fail(cudaError_t e, const char *what) {
    return fail(HTB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

namespace {
std::mutex g_option_mutex;
std::map<std::string, int64_t> g_options = {
    {"block_rows", 0}, {"piece_cols", 16}, {"stage_bytes", 24576}, {"cseg_bytes", 4096}, {"ring_stages", 2}, {"reduce_ring_stages", 3}, {"evict_first", 1}, {"upload_chunk_mb", 256}, {"m_ring_stages", 0}, {"m_b_ring_log2", 3}, {"m_reduce_ring_stages", 0}, {"mrhs_min", 0}, {"pack_generate_dense", 0}, {"sort_units", 1}, {"reduce_blocks_per_cta", 0}, {"m_reduce_warps", 24}, {"m_pad", 4}, {"m_b_producers", 3}, {"m_stage_input", 1}, {"m_small_runs", 1}, {"m_reduce_split", 1}, {"fused_symmetric", 1}, {"dist_p2p", 1}, {"target_block_rows", 0}, {"zero_copy", 1}, {"tail_split", 1}, {"cta_slots", 0}, {"pdl", 1}, {"aca_fma_axpy", 0}, {"aca_rank_guess", 16}, {"aca_dots", 0}, {"upload_headers_only", 1}, {"m_near_field", 0}, {"m_b_global", 1}, {"m_fast_tall", 1}, {"m_nf_rows", 0}};

int64_t option(const char *key) {
    std::lock_guard<std::mutex> lock(g_option_mutex);
    return g_options.at(key);
}
} // namespace

int64_t option_value(const char *key) { return option(key); }

#define HTB_CUDA(call)                       \
    do {                                     \
        cudaError_t e__ = (call);            \
        if (e__ != cudaSuccess)              \
            return cuda_fail(e__, #call);    \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok  = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess)
            ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

template <typename V, typename A>
static int upload_vector(const std::vector<V, A> &v, const V **dptr, std::vector<void *> &owned) {
    *dptr = nullptr;
    if (v.empty())
        return HTB_OK;
    void *p = nullptr;
    HTB_CUDA(cudaMalloc(&p, v.size() * sizeof(V)));
    owned.push_back(p);
    HTB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    *dptr = static_cast<const V *>(p);
    return HTB_OK;
}

static int upload_store(htb_operator *h, const Packer &pk) {
    const bool timing = std::getenv("HTB_PACK_TIMING") != nullptr; // development aid: seconds of every phase on stderr
    auto t_last       = std::chrono::steady_clock::now();
    auto lap          = [&](const char *what) {
        if (!timing)
            return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[htb upload] %-22s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    const size_t chunk = static_cast<size_t>(std::max<int64_t>(1, option("upload_chunk_mb"))) << 20;
    // two pinned buffers: the CPU packs block streams into one while the other is in flight to the device
    const bool headers_only = pk.all_on_device && option("upload_headers_only") != 0;
    size_t need = 0, largest_block = 0;
    for (int s = 0; s < 2; s++) {
        need = std::max<size_t>(need, headers_only ? pk.header_offset(s, pk.side[s].stages.size()) : pk.side[s].stream_bytes);
        if (headers_only)
            for (const BlockDesc &bd : pk.side[s].blocks)
                largest_block = std::max<size_t>(largest_block, pk.header_offset(s, bd.first_stage + bd.n_stages) - pk.header_offset(s, bd.first_stage));
    }
    // (page-locking costs ~0.4 s per GB: the headers of a device-assembled store travel through 2 x 32 MB)
    const size_t buf_bytes = std::min(need, headers_only ? std::max<size_t>(largest_block, size_t(32) << 20) : chunk);
    char *pinned[2]        = {nullptr, nullptr};
    cudaEvent_t done[2]    = {nullptr, nullptr};
    int status             = HTB_OK;
    auto cleanup           = [&]() {
        for (int i = 0; i < 2; i++) {
            if (pinned[i])
                cudaFreeHost(pinned[i]);
            if (done[i])
                cudaEventDestroy(done[i]);
        }
    };
    if (buf_bytes) {
        for (int i = 0; i < 2; i++) {
            cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&pinned[i]), buf_bytes);
            if (e == cudaSuccess)
                e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
            if (e != cudaSuccess) {
                cleanup();
                return cuda_fail(e, "pinned upload buffers");
            }
        }
    }
    lap("pinned buffers");
    for (int s = 0; s < 2 && status == HTB_OK; s++) {
        const SideLayout &sl = pk.side[s];
        SideDevice &sd       = h->side[s];
        sd.n                 = sl.n;
        sd.n_blocks          = static_cast<int>(sl.blocks.size());
        sd.n_combine         = static_cast<int>(sl.combine.size());
        sd.any_twice         = sl.any_twice;
        sd.cs_base           = sl.cs_base;
        h->host_blocks[s]    = sl.blocks;
        h->host_order[s]     = sl.order;
        if ((status = upload_vector(sl.blocks, &sd.blocks, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.stages, &sd.stages, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.order, &sd.order, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.combine, &sd.combine, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.combine_dst, &sd.combine_dst, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.munits, &sd.munits, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.combine_m, &sd.combine_m, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.aux_reduce, &sd.aux_reduce, h->owned)) != HTB_OK)
            break;
        if ((status = upload_vector(sl.aux_apply, &sd.aux_apply, h->owned)) != HTB_OK)
            break;
        h->descriptor_bytes += 2 * sl.aux_reduce.size();
        sd.n_combine_m = static_cast<int>(sl.combine_m.size());
        h->descriptor_bytes += sl.munits.size() * sizeof(MUnit) + sl.combine_m.size() * sizeof(CombineEntry);
        h->descriptor_bytes += sl.blocks.size() * sizeof(BlockDesc) + sl.stages.size() * sizeof(StageDesc) + sl.order.size() * 4 + sl.combine.size() * sizeof(CombineEntry) + sl.combine_dst.size() * sizeof(CombineDst);
        lap("descriptor tables");
        if (sl.stream_bytes == 0)
            continue;
        void *dstream = nullptr;
        cudaError_t e = cudaMalloc(&dstream, sl.stream_bytes);
        if (e != cudaSuccess) {
            status = cuda_fail(e, "cudaMalloc(leaf store)");
            break;
        }
        h->owned.push_back(dstream);
        sd.stream       = static_cast<const unsigned char *>(dstream);
        sd.stream_bytes = sl.stream_bytes;
        h->store_bytes += sl.stream_bytes;
        h->side_stream_bytes[s] = sl.stream_bytes;
        const int nb = sd.n_blocks;
        if (headers_only) {
            // Device assembly: every panel is generated / copied on the device, the stream is zeros + the stage headers. Only the
            // headers travel (16 + 16 n_units bytes per stage, ~2 % of the stream), packed back to back, and a kernel puts them in place.
            const uint64_t total_hdr = pk.header_offset(s, sl.stages.size());
            void *d_compact = nullptr, *d_off = nullptr;
            e = cudaMemsetAsync(dstream, 0, sl.stream_bytes, h->own_stream);
            if (e == cudaSuccess)
                e = cudaMalloc(&d_compact, std::max<uint64_t>(16, total_hdr));
            if (e == cudaSuccess)
                e = cudaMalloc(&d_off, (sl.stages.size() + 1) * sizeof(uint64_t));
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(d_off, pk.header_offsets(s).data(), (sl.stages.size() + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->own_stream);
            int b0 = 0, turn = 0;
            while (e == cudaSuccess && b0 < nb) {
                auto hoff = [&](int b) { return b < nb ? pk.header_offset(s, sl.blocks[b].first_stage) : total_hdr; };
                int b1 = b0 + 1;
                while (b1 < nb && hoff(b1 + 1) - hoff(b0) <= buf_bytes)
                    b1++;
                const uint64_t off = hoff(b0), bytes = hoff(b1) - off;
                if (bytes > buf_bytes) {
                    status = fail(HTB_ERR_ALLOC, "block headers larger than the upload buffer");
                    break;
                }
                if (bytes) {
                    cudaEventSynchronize(done[turn]);
                    pk.fill_headers(s, b0, b1, pinned[turn]);
                    e = cudaMemcpyAsync(static_cast<char *>(d_compact) + off, pinned[turn], bytes, cudaMemcpyHostToDevice, h->own_stream);
                    if (e == cudaSuccess)
                        e = cudaEventRecord(done[turn], h->own_stream);
                    turn ^= 1;
                }
                b0 = b1;
            }
            if (e == cudaSuccess && status == HTB_OK)
                e = launch_scatter_headers(sd.stages, static_cast<const unsigned long long *>(d_off), static_cast<long long>(sl.stages.size()), static_cast<const unsigned char *>(d_compact), static_cast<unsigned char *>(dstream), h->own_stream);
            if (e == cudaSuccess)
                e = cudaStreamSynchronize(h->own_stream);
            if (d_compact)
                cudaFree(d_compact);
            if (d_off)
                cudaFree(d_off);
            if (e != cudaSuccess && status == HTB_OK)
                status = cuda_fail(e, "upload of the stage headers");
            if (status != HTB_OK)
                break;
            h->launches++;
            lap("stream (headers only)");
            continue;
        }
        // batches of consecutive blocks whose streams fit one pinned buffer
        int b0 = 0, turn = 0;
        while (b0 < nb) {
            int b1 = b0 + 1;
            while (b1 < nb && pk.block_offset(s, b1 + 1) - pk.block_offset(s, b0) <= buf_bytes)
                b1++;
            const uint64_t off = pk.block_offset(s, b0), bytes = pk.block_offset(s, b1) - off;
            if (bytes > buf_bytes) { // cannot happen: a block stream is bounded by the chunk size check below
                status = fail(HTB_ERR_ALLOC, "block stream larger than the upload buffer");
                break;
            }
            if (bytes) {
                cudaEventSynchronize(done[turn]);
                pk.fill(s, b0, b1, pinned[turn]);
                e = cudaMemcpyAsync(static_cast<char *>(dstream) + off, pinned[turn], bytes, cudaMemcpyHostToDevice, h->own_stream);
                if (e == cudaSuccess)
                    e = cudaEventRecord(done[turn], h->own_stream);
                if (e != cudaSuccess) {
                    status = cuda_fail(e, "upload of the leaf store");
                    break;
                }
                turn ^= 1;
            }
            b0 = b1;
        }
        lap("stream");
    }
    cudaError_t e = cudaStreamSynchronize(h->own_stream);
    cleanup();
    lap("cleanup");
    if (status == HTB_OK && e != cudaSuccess)
        status = cuda_fail(e, "upload of the leaf store");
    return status;
}

// ---- products ---------------------------------------------------------------------------------------

template <typename T>
static int run_product(htb_operator *h, char trans, T alpha, const T *in, T beta, T *out, int stride, const DistSplit *split) {
    const char sym = h->symmetry;
    const bool twice = sym != 'N' && (h->side[0].any_twice || h->side[1].any_twice);
    const int D      = h->row_offset - h->col_offset;
    T *T1 = static_cast<T *>(h->d_scratch);
    T *T2 = T1 + h->scratch_elems;
    cudaStream_t st = h->stream;
    const bool is_complex = h->dtype == HTB_COMPLEX_DOUBLE;
    T one;
    std::memset(&one, 0, sizeof(T));
    *reinterpret_cast<double *>(&one) = 1.0;

    // every launch goes through here: counts it and, when profiling, brackets it with events on its stream
    auto timed = [&](int kind, auto &&launch, const char *what) -> int {
        htb_operator::TimedLaunch tl{nullptr, nullptr, kind};
        if (h->profiling) {
            cudaEventCreate(&tl.start);
            cudaEventCreate(&tl.stop);
            cudaEventRecord(tl.start, st);
        }
        cudaError_t e = launch();
        if (h->profiling) {
            cudaEventRecord(tl.stop, st);
            h->timed.push_back(tl);
        }
        if (e != cudaSuccess)
            return cuda_fail(e, what);
        h->launches++;
        return HTB_OK;
    };
#define count(expr, what) timed(std::strstr(what, "reduce") ? HTB_PASS_REDUCE : (std::strstr(what, "combine") ? HTB_PASS_COMBINE : HTB_PASS_APPLY), [&]() { return (expr); }, what)
    int rc;
    if (trans == 'N') {
        // direction 0 producers: t = V x for every low-rank leaf (side 1 holds the V^T panels) and the x slices of
        // the dense leaves, written into the c-stream of side 0 (or into partials when a piece has several chunks)
        PassArgs<T> r1;
        r1.in = in, r1.in_len = h->nb_cols, r1.scratch = T1, r1.stride = stride;
        if (h->side[1].stream && !split) {
            if ((rc = count(launch_reduce<T>(h->side[1], h->launch_cfg, r1, st), "reduce(V)")) != HTB_OK)
                return rc;
        } else if (h->side[1].stream && split->flags) {
            // distributed, x gathered through peer memory: one launch, own-partition blocks first, the others wait in
            // the kernel for the arrival flag of the rank that owns their slice
            SideDevice part = h->side[1];
            part.order = split->order_all, part.n_blocks = split->n_all;
            PassArgs<T> rw = r1;
            rw.wait_flags = split->flags, rw.wait_owner = split->owner, rw.wait_epoch = split->epoch;
            if (part.n_blocks && (rc = count(launch_reduce<T>(part, h->launch_cfg, rw, st), "reduce(V, peer gather)")) != HTB_OK)
                return rc;
        } else if (h->side[1].stream) {
            // distributed: source blocks inside the rank's own partition first, they overlap the allgather of x
            SideDevice part = h->side[1];
            part.order = split->order_local, part.n_blocks = split->n_local;
            if (part.n_blocks && (rc = count(launch_reduce<T>(part, h->launch_cfg, r1, st), "reduce(V, local)")) != HTB_OK)
                return rc;
        }
        const bool fuse = twice && h->fused_symmetric;
        if (twice && !fuse) {
            // second application of the leaves stored once under symmetry (add_hmatrix_vector_product.hpp:154-163):
            // direction 1 producers restricted to those leaves, t' = op(U)^T x[target], z = op(A)^T x[target], op = conj for 'H'
            PassArgs<T> r0;
            r0.in = in, r0.in_len = h->nb_cols, r0.in_shift = D, r0.scratch = T2, r0.twice_only = 1, r0.conj = (sym == 'H' && is_complex), r0.stride = stride;
            if ((rc = count(launch_reduce<T>(h->side[0], h->launch_cfg, r0, st), "reduce(U, twice)")) != HTB_OK)
                return rc;
            if (h->side[1].n_combine && (rc = count(launch_combine<T>(h->side[1], T2, 1, st), "combine(U, twice)")) != HTB_OK)
                return rc;
        }
        if (h->side[1].stream && split && !split->flags) {
            cudaError_t we = cudaStreamWaitEvent(st, split->gather_done, 0);
            if (we != cudaSuccess)
                return cuda_fail(we, "cudaStreamWaitEvent(allgather)");
            SideDevice part = h->side[1];
            part.order = split->order_remote, part.n_blocks = split->n_remote;
            if (part.n_blocks && (rc = count(launch_reduce<T>(part, h->launch_cfg, r1, st), "reduce(V, remote)")) != HTB_OK)
                return rc;
        }
        if (h->side[1].stream && h->side[0].n_combine && (rc = count(launch_combine<T>(h->side[0], T1, 0, st), "combine(V)")) != HTB_OK)
            return rc;
        // y = beta y + alpha (U t + A x), rows owned by one CTA each. Under symmetric storage the same pass ALSO reduces the
        // stored-once leaves against x[target] (fused): side 0 is streamed once for both applications
        PassArgs<T> a0;
        a0.out = out, a0.out_len = h->nb_rows, a0.scratch = T1, a0.alpha = alpha, a0.beta = beta, a0.stride = stride;
        if (fuse)
            a0.fused = 1, a0.in = in, a0.in_len = h->nb_cols, a0.in_shift = D, a0.scratch2 = T2, a0.conj2 = (sym == 'H' && is_complex);
        if ((rc = count(launch_apply<T>(h->side[0], h->launch_cfg, a0, st), "apply(U, A)")) != HTB_OK)
            return rc;
        if (fuse && h->side[1].n_combine && (rc = count(launch_combine<T>(h->side[1], T2, 1, st), "combine(U, twice)")) != HTB_OK)
            return rc;
        if (twice) {
            PassArgs<T> a1;
            a1.out = out, a1.out_len = h->nb_rows, a1.out_shift = -D, a1.scratch = T2, a1.alpha = alpha, a1.beta = one, a1.twice_only = 1, a1.conj = (sym == 'H' && is_complex), a1.stride = stride;
            if ((rc = count(launch_apply<T>(h->side[1], h->launch_cfg, a1, st), "apply(V^T, twice)")) != HTB_OK)
                return rc;
        }
    } else {
        const int conj = (trans == 'C' && is_complex) ? 1 : 0;
        PassArgs<T> r0;
        r0.in = in, r0.in_len = h->nb_rows, r0.scratch = T1, r0.conj = conj, r0.stride = stride;
        if (h->side[0].stream) {
            if ((rc = count(launch_reduce<T>(h->side[0], h->launch_cfg, r0, st), "reduce(U, A)")) != HTB_OK)
                return rc;
            if (h->side[1].n_combine && (rc = count(launch_combine<T>(h->side[1], T1, 0, st), "combine(U)")) != HTB_OK)
                return rc;
        }
        const bool fuse = twice && h->fused_symmetric;
        if (twice && !fuse) {
            PassArgs<T> r1;
            r1.in = in, r1.in_len = h->nb_rows, r1.in_shift = -D, r1.scratch = T2, r1.twice_only = 1, r1.stride = stride;
            if ((rc = count(launch_reduce<T>(h->side[1], h->launch_cfg, r1, st), "reduce(V, twice)")) != HTB_OK)
                return rc;
            if (h->side[0].n_combine && (rc = count(launch_combine<T>(h->side[0], T2, 1, st), "combine(V, twice)")) != HTB_OK)
                return rc;
        }
        PassArgs<T> a1;
        a1.out = out, a1.out_len = h->nb_cols, a1.scratch = T1, a1.alpha = alpha, a1.beta = beta, a1.conj = conj, a1.stride = stride;
        if (fuse)
            a1.fused = 1, a1.in = in, a1.in_len = h->nb_rows, a1.in_shift = -D, a1.scratch2 = T2, a1.conj2 = 0;
        if ((rc = count(launch_apply<T>(h->side[1], h->launch_cfg, a1, st), "apply(V^T)")) != HTB_OK)
            return rc;
        if (fuse && h->side[0].n_combine && (rc = count(launch_combine<T>(h->side[0], T2, 1, st), "combine(V, twice)")) != HTB_OK)
            return rc;
        if (twice) {
            PassArgs<T> a0;
            a0.out = out, a0.out_len = h->nb_cols, a0.out_shift = D, a0.scratch = T2, a0.alpha = alpha, a0.beta = one, a0.twice_only = 1, a0.stride = stride;
            if ((rc = count(launch_apply<T>(h->side[0], h->launch_cfg, a0, st), "apply(U, A, twice)")) != HTB_OK)
                return rc;
        }
    }
    return HTB_OK;
#undef count
}

// ---- multi-RHS products on the FP64 tensor cores (double, mu >= mrhs_min) -------------------------------------------
static int ensure_mscratch(htb_operator *h, int vs) {
    if (h->d_mscratch && h->mscratch_vs >= vs)
        return HTB_OK;
    if (h->d_mscratch)
        cudaFree(h->d_mscratch);
    h->d_mscratch  = nullptr;
    h->mscratch_vs = 0;
    const size_t bytes = std::max<size_t>(1, h->mscratch_elems) * (vs + 8) * sizeof(double) * (h->needs_second_copy ? 2 : 1);
    cudaError_t e      = cudaMalloc(&h->d_mscratch, bytes);
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaMalloc(multi-RHS scratch)");
    h->mscratch_vs = vs;
    return HTB_OK;
}

// Device copy of the multi-RHS near field, built from the main stream of side 0 the first time a multi-RHS product 'N' runs.
// A failure (e.g. out of memory for the second copy of the dense coefficients) is not an error: the product keeps the dense
// columns in the runs of the main stream.
static void ensure_near_field(htb_operator *h) {
    if (h->nf_ready || h->nf_failed || !h->nf_host)
        return;
    const NearFieldLayout &nf = *h->nf_host;
    std::vector<void *> mine;
    void *d_tasks = nullptr, *d_hdr = nullptr, *d_off = nullptr, *d_stream = nullptr;
    auto up = [&](const void *src, size_t bytes, void **dst) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(16, bytes));
        if (e == cudaSuccess && bytes)
            e = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, h->own_stream);
        return e;
    };
    void *d_blocks = nullptr, *d_stages = nullptr, *d_order = nullptr, *d_aux = nullptr;
    cudaError_t e = up(nf.blocks.data(), nf.blocks.size() * sizeof(BlockDesc), &d_blocks);
    if (e == cudaSuccess)
        e = up(nf.stages.data(), nf.stages.size() * sizeof(StageDesc), &d_stages);
    if (e == cudaSuccess)
        e = up(nf.order.data(), nf.order.size() * sizeof(uint32_t), &d_order);
    if (e == cudaSuccess)
        e = up(nf.aux_apply.data(), nf.aux_apply.size(), &d_aux);
    if (e == cudaSuccess)
        e = up(nf.tasks.data(), nf.tasks.size() * sizeof(NfTask), &d_tasks);
    if (e == cudaSuccess)
        e = up(nf.headers.data(), nf.headers.size(), &d_hdr);
    if (e == cudaSuccess)
        e = up(nf.hdr_off.data(), nf.hdr_off.size() * sizeof(uint64_t), &d_off);
    if (e == cudaSuccess)
        e = cudaMalloc(&d_stream, std::max<uint64_t>(16, nf.stream_bytes));
    if (e == cudaSuccess)
        e = cudaMemsetAsync(d_stream, 0, nf.stream_bytes, h->own_stream);
    if (e == cudaSuccess)
        e = launch_scatter_headers(static_cast<const StageDesc *>(d_stages), static_cast<const unsigned long long *>(d_off), static_cast<long long>(nf.stages.size()), static_cast<const unsigned char *>(d_hdr),
                                   static_cast<unsigned char *>(d_stream), h->own_stream);
    if (e == cudaSuccess)
        e = launch_nf_copy(static_cast<const NfTask *>(d_tasks), static_cast<long long>(nf.tasks.size()), h->side[0].stream, static_cast<unsigned char *>(d_stream), static_cast<int>(h->esize), h->own_stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(h->own_stream);
    for (void *p : {d_tasks, d_hdr, d_off})
        if (p)
            cudaFree(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        for (void *p : {d_blocks, d_stages, d_order, d_aux, d_stream})
            if (p)
                cudaFree(p);
        h->nf_failed = true;
        h->nf_host.reset();
        return;
    }
    for (void *p : {d_blocks, d_stages, d_order, d_aux, d_stream})
        h->owned.push_back(p);
    h->nf.blocks    = static_cast<const BlockDesc *>(d_blocks);
    h->nf.stages    = static_cast<const StageDesc *>(d_stages);
    h->nf.order     = static_cast<const uint32_t *>(d_order);
    h->nf.stream    = static_cast<const unsigned char *>(d_stream);
    h->nf.aux_apply = static_cast<const unsigned char *>(d_aux);
    h->nf.n_blocks  = static_cast<int>(nf.blocks.size());
    h->nf.stream_bytes = nf.stream_bytes;
    h->nf.n            = h->side[0].n;
    h->nf_bytes        = nf.stream_bytes + nf.aux_apply.size();
    h->store_bytes += nf.stream_bytes;
    h->descriptor_bytes += nf.aux_apply.size() + nf.blocks.size() * sizeof(BlockDesc) + nf.stages.size() * sizeof(StageDesc);
    h->launches += 2;
    h->nf_ready = true;
    h->nf_host.reset();
}

static int run_product_m(htb_operator *h, char trans, const double *alpha, const double *in, const double *beta, double *out, int mu) {
    const char sym   = h->symmetry;
    const bool twice = sym != 'N' && (h->side[0].any_twice || h->side[1].any_twice);
    const int D      = h->row_offset - h->col_offset;
    const bool cplx  = h->dtype == HTB_COMPLEX_DOUBLE;
    const int W      = cplx ? 2 : 1; // doubles per entry: the kernels address complex matrices through their real view
    const int group  = 64 / W;       // right-hand sides per pass
    cudaStream_t st  = h->stream;
    int rc;
    auto timed = [&](int kind, auto &&launch, const char *what) -> int {
        htb_operator::TimedLaunch tl{nullptr, nullptr, kind};
        if (h->profiling) {
            cudaEventCreate(&tl.start);
            cudaEventCreate(&tl.stop);
            cudaEventRecord(tl.start, st);
        }
        cudaError_t e = launch();
        if (h->profiling) {
            cudaEventRecord(tl.stop, st);
            h->timed.push_back(tl);
        }
        if (e != cudaSuccess)
            return cuda_fail(e, what);
        h->launches++;
        return HTB_OK;
    };
    if (trans == 'N' && option("m_near_field") != 0)
        ensure_near_field(h);
    for (int col0 = 0; col0 < mu; col0 += group) {
        const int mc = std::min(group, mu - col0) * W, vs = (mc + 7) & ~7;
        if ((rc = ensure_mscratch(h, (std::min(group, mu) * W + 7) & ~7)) != HTB_OK)
            return rc;
        double *M1 = static_cast<double *>(h->d_mscratch);
        double *M2 = M1 + h->mscratch_elems * static_cast<size_t>(h->mscratch_vs + 8); // (allocated with the padded stride: any vsp <= vs + 8 fits)
        MArgs base;
        // vsp: vector stride of the scratch = row stride of APPLY_M's B ring and of REDUCE_M's X block; vs + 4 keeps the DMMA
        // fragment loads free of bank conflicts (mkernels.cuh)
        const int pad = h->launch_cfg.m_pad;
        base.ld_in = mu * W, base.ld_out = mu * W, base.col0 = col0 * W, base.col0_in = col0 * W, base.mc = mc, base.vs = vs, base.vsp = vs + pad, base.cplx = cplx ? 1 : 0;
        base.b_global   = option("m_b_global") != 0;
        base.fast_tall  = option("m_fast_tall") != 0;
        base.small_runs = option("m_small_runs") != 0, base.reduce_split = static_cast<int>(std::min<int64_t>(8, std::max<int64_t>(1, option("m_reduce_split"))));
        // the group's columns of the input, copied once into rows one B-ring row apart (zero padded): the input rows of a dense
        // leaf then reach the B ring with one bulk copy whatever mu is, and every row is 16 B aligned
        const long long in_rows_all = trans == 'N' ? h->nb_cols : h->nb_rows;
        const double *in_g          = in;
        if (option("m_stage_input") != 0 && (base.ld_in != base.vsp || col0 != 0)) {
            const size_t need = static_cast<size_t>(std::max<long long>(1, in_rows_all)) * base.vsp * sizeof(double);
            if (need > h->mstage_cap) {
                if (h->d_mstage)
                    cudaFree(h->d_mstage);
                h->d_mstage = nullptr, h->mstage_cap = 0;
                cudaError_t e = cudaMalloc(&h->d_mstage, need);
                if (e != cudaSuccess)
                    return cuda_fail(e, "cudaMalloc(multi-RHS input staging)");
                h->mstage_cap = need;
            }
            if ((rc = timed(HTB_PASS_OTHER, [&]() { return launch_stage_group(in, in_rows_all, base.ld_in, base.col0_in, mc, static_cast<double *>(h->d_mstage), base.vsp, st); }, "stage_group")) != HTB_OK)
                return rc;
            in_g = static_cast<const double *>(h->d_mstage), base.ld_in = base.vsp, base.col0_in = 0;
        }
        base.alpha = alpha[0], base.alpha_im = cplx ? alpha[1] : 0.;
        // ps: side streamed by REDUCE_M (producers), cs: side streamed by APPLY_M (consumers)
        auto direction = [&](int cs, double *M, int in_shift, long long in_rows, int out_shift, long long out_rows, double b_re, double b_im, int twice_only, int conj) -> int {
            const int ps = 1 - cs;
            MArgs r      = base;
            r.in = in_g, r.in_rows = in_rows, r.in_shift = in_shift, r.mscratch = M, r.twice_only = twice_only, r.conj = conj;
            int rc2;
            if (h->side[ps].stream && (rc2 = timed(HTB_PASS_REDUCE, [&]() { return launch_reduce_m(h->side[ps], h->launch_cfg, r, st); }, "reduce_m")) != HTB_OK)
                return rc2;
            if (h->side[cs].n_combine_m && (rc2 = timed(HTB_PASS_COMBINE, [&]() { return launch_combine_m(h->side[cs], M, vs, base.vsp, twice_only, st); }, "combine_m")) != HTB_OK)
                return rc2;
            MArgs ap = r;
            ap.out = out, ap.out_rows = out_rows, ap.out_shift = out_shift, ap.beta = b_re, ap.beta_im = b_im;
            // first application of 'N': the dense leaves come from the near-field panels (one full-height panel per target block)
            const bool near_field = cs == 0 && !twice_only && !conj && h->nf_ready && option("m_near_field") != 0;
            ap.skip_dense         = near_field ? 1 : 0;
            if ((rc2 = timed(HTB_PASS_APPLY, [&]() { return launch_apply_m(h->side[cs], h->launch_cfg, ap, st); }, "apply_m")) != HTB_OK || !near_field)
                return rc2;
            MArgs nfa      = ap;
            nfa.skip_dense = 0, nfa.beta = 1., nfa.beta_im = 0.;
            return timed(std::getenv("HTB_NF_AS_OTHER") ? HTB_PASS_OTHER : HTB_PASS_APPLY, [&]() { return launch_apply_m(h->nf, h->launch_cfg, nfa, st); }, "apply_m(near field)");
        };
        const double b_re = beta[0], b_im = cplx ? beta[1] : 0.;
        const int herm = (sym == 'H' && cplx) ? 1 : 0; // second application of a Hermitian leaf stored once: conjugate-transposed
        if (trans == 'N') {
            if ((rc = direction(0, M1, 0, h->nb_cols, 0, h->nb_rows, b_re, b_im, 0, 0)) != HTB_OK)
                return rc;
            if (twice && (rc = direction(1, M2, D, h->nb_cols, -D, h->nb_rows, 1.0, 0.0, 1, herm)) != HTB_OK)
                return rc;
        } else {
            const int conj = (trans == 'C' && cplx) ? 1 : 0;
            if ((rc = direction(1, M1, 0, h->nb_rows, 0, h->nb_cols, b_re, b_im, 0, conj)) != HTB_OK)
                return rc;
            if (twice && (rc = direction(0, M2, -D, h->nb_rows, D, h->nb_cols, 1.0, 0.0, 1, 0)) != HTB_OK)
                return rc;
        }
    }
    return HTB_OK;
}

static int check_trans(const htb_operator *h, char trans) {
    if (trans != 'N' && trans != 'T' && trans != 'C')
        return fail(HTB_ERR_INVALID, std::string("unknown trans '") + trans + "'");
    // same condition and wording as add_hmatrix_vector_product.hpp:112-115
    if ((trans == 'T' && h->symmetry == 'H') || (trans == 'C' && h->symmetry == 'S'))
        return fail(HTB_ERR_UNSUPPORTED, std::string("Operation is not supported (trans=") + trans + " with " + h->symmetry + ")");
    return HTB_OK;
}

// device pointers, mu right-hand sides row-major
int product_device(htb_operator *h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, const DistSplit *split) {
    int rc = check_trans(h, trans);
    if (rc != HTB_OK)
        return rc;
    // From how many right-hand sides the tensor-core path wins over a loop of single-RHS products (measured, N = 1e6: the
    // multi-RHS passes cost ~5 single-RHS products whatever mu <= 64 is): option mrhs_min, 0 = automatic
    const int mrhs_auto = h->dtype == HTB_COMPLEX_DOUBLE ? 4 : 5;
    const int mrhs_min  = option("mrhs_min") > 0 ? static_cast<int>(option("mrhs_min")) : mrhs_auto;
    if (h->m_path_ok && mu >= mrhs_min) {
        // tensor-core path: one pass over all columns, so a distributed product first waits for the gather of x
        if (split && split->flags)
            HTB_CUDA(launch_wait_flags(split->flags, split->world, split->epoch, h->stream));
        else if (split && split->gather_done)
            HTB_CUDA(cudaStreamWaitEvent(h->stream, split->gather_done, 0));
        return run_product_m(h, trans, static_cast<const double *>(alpha), static_cast<const double *>(in), static_cast<const double *>(beta), static_cast<double *>(out), mu);
    }
    for (int c = 0; c < mu; c++) {
        if (h->dtype == HTB_DOUBLE)
            rc = run_product<double>(h, trans, *static_cast<const double *>(alpha), static_cast<const double *>(in) + c, *static_cast<const double *>(beta), static_cast<double *>(out) + c, mu, c == 0 ? split : nullptr);
        else
            rc = run_product<cplx>(h, trans, *static_cast<const cplx *>(alpha), static_cast<const cplx *>(in) + c, *static_cast<const cplx *>(beta), static_cast<cplx *>(out) + c, mu, c == 0 ? split : nullptr);
        if (rc != HTB_OK)
            return rc;
    }
    return HTB_OK;
}

int ensure_staging(htb_operator *h, size_t in_bytes, size_t out_bytes) {
    auto grow = [&](void **dev, void **host, size_t *cap, size_t need) -> int {
        if (need <= *cap)
            return HTB_OK;
        if (*dev)
            cudaFree(*dev);
        if (*host)
            cudaFreeHost(*host);
        *dev = *host = nullptr;
        *cap         = 0;
        HTB_CUDA(cudaMalloc(dev, need));
        HTB_CUDA(cudaMallocHost(host, need));
        *cap = need;
        return HTB_OK;
    };
    int rc = grow(&h->d_in, &h->h_in, &h->in_cap, in_bytes);
    if (rc != HTB_OK)
        return rc;
    return grow(&h->d_out, &h->h_out, &h->out_cap, out_bytes);
}

constexpr size_t kStageChunk = size_t(1) << 20;

// Page-locked host memory (cudaMallocHost / cudaHostRegister / htb_host_register) is copied by the DMA engines directly
// (both ends of the span are checked: a vector that starts inside a registered range but runs past its end must be staged)
static bool is_pinned_host_byte(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}
bool is_pinned_host(const void *p, size_t bytes) {
    return is_pinned_host_byte(p) && (bytes <= 1 || is_pinned_host_byte(static_cast<const char *>(p) + bytes - 1));
}

// pageable -> pinned copy of one chunk, split over the OpenMP threads of the caller's process
void host_copy(char *dst, const char *src, size_t n) {
    constexpr size_t kPiece = size_t(128) << 10;
    const long long pieces  = static_cast<long long>((n + kPiece - 1) / kPiece);
#pragma omp parallel for schedule(static) if (pieces >= 4)
    for (long long i = 0; i < pieces; i++) {
        const size_t off = static_cast<size_t>(i) * kPiece;
        std::memcpy(dst + off, src + off, std::min(kPiece, n - off));
    }
}

int staged_h2d(htb_operator *, void *dev, void *pinned, const void *host, size_t bytes, cudaStream_t st) {
    if (is_pinned_host(host, bytes)) {
        HTB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
        return HTB_OK;
    }
    for (size_t off = 0; off < bytes; off += kStageChunk) {
        const size_t n = std::min(kStageChunk, bytes - off);
        host_copy(static_cast<char *>(pinned) + off, static_cast<const char *>(host) + off, n);
        HTB_CUDA(cudaMemcpyAsync(static_cast<char *>(dev) + off, static_cast<char *>(pinned) + off, n, cudaMemcpyHostToDevice, st));
    }
    return HTB_OK;
}

int staged_d2h(htb_operator *h, void *host, void *pinned, const void *dev, size_t bytes, cudaStream_t st) {
    if (is_pinned_host(host, bytes)) {
        HTB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
        HTB_CUDA(cudaStreamSynchronize(st));
        return HTB_OK;
    }
    const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
    while (h->chunk_events.size() < n_chunks) {
        cudaEvent_t ev;
        HTB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->chunk_events.push_back(ev);
    }
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t off = c * kStageChunk, n = std::min(kStageChunk, bytes - off);
        HTB_CUDA(cudaMemcpyAsync(static_cast<char *>(pinned) + off, static_cast<const char *>(dev) + off, n, cudaMemcpyDeviceToHost, st));
        HTB_CUDA(cudaEventRecord(h->chunk_events[c], st));
    }
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t off = c * kStageChunk, n = std::min(kStageChunk, bytes - off);
        HTB_CUDA(cudaEventSynchronize(h->chunk_events[c]));
        host_copy(static_cast<char *>(host) + off, static_cast<char *>(pinned) + off, n);
    }
    HTB_CUDA(cudaStreamSynchronize(st));
    return HTB_OK;
}

static bool beta_is_zero(const htb_operator *h, const void *beta) {
    const double *b = static_cast<const double *>(beta);
    return b[0] == 0. && (h->dtype == HTB_DOUBLE || b[1] == 0.);
}

// Device address of a page-locked, mapped host buffer (cudaHostRegister / htb_host_register / cudaMallocHost), or nullptr
void *mapped_device_pointer(const void *host, size_t bytes) {
    if (!is_pinned_host(host, bytes))
        return nullptr;
    void *dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, const_cast<void *>(host), 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return dev;
}

// host pointers: stage through pinned buffers, H2D, product, D2H, synchronise
static int product_host(htb_operator *h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) {
    int rc = check_trans(h, trans);
    if (rc != HTB_OK)
        return rc;
    // Zero copy (single RHS, page-locked mapped buffers): every entry of x is read exactly once per REDUCE block prologue
    // and every entry of y is written exactly once by the APPLY epilogue, so the kernels address the HOST vectors
    // directly over PCIe — no H2D / D2H phase before and after the product, the 8 B per row ride along with the
    // 20 kB per row of coefficients streamed from HBM.
    if (mu == 1 && option("zero_copy") != 0) {
        void *din  = mapped_device_pointer(in, static_cast<size_t>(trans == 'N' ? h->nb_cols : h->nb_rows) * h->esize);
        void *dout = mapped_device_pointer(out, static_cast<size_t>(trans == 'N' ? h->nb_rows : h->nb_cols) * h->esize);
        if (din && dout) {
            if ((rc = product_device(h, trans, alpha, din, beta, dout, 1)) != HTB_OK)
                return rc;
            HTB_CUDA(cudaStreamSynchronize(h->stream));
            return HTB_OK;
        }
    }
    const size_t ni = trans == 'N' ? h->nb_cols : h->nb_rows, no = trans == 'N' ? h->nb_rows : h->nb_cols;
    const size_t in_bytes = ni * mu * h->esize, out_bytes = no * mu * h->esize;
    if ((rc = ensure_staging(h, in_bytes, out_bytes)) != HTB_OK)
        return rc;
    cudaStream_t st = h->stream;
    if ((rc = staged_h2d(h, h->d_in, h->h_in, in, in_bytes, st)) != HTB_OK)
        return rc;
    if (!beta_is_zero(h, beta) && (rc = staged_h2d(h, h->d_out, h->h_out, out, out_bytes, st)) != HTB_OK)
        return rc;
    if ((rc = product_device(h, trans, alpha, h->d_in, beta, h->d_out, mu)) != HTB_OK)
        return rc;
    return staged_d2h(h, out, h->h_out, h->d_out, out_bytes, st);
}

} // namespace htb

using namespace htb;

extern "C" {

const char *htb_last_error(void) { return g_last_error.c_str(); }

int htb_device_count(int *count) {
    if (!count)
        return fail(HTB_ERR_INVALID, "null argument");
    *count = 0;
    HTB_CUDA(cudaGetDeviceCount(count));
    return HTB_OK;
}

int htb_host_register(void *ptr, size_t bytes) {
    if (!ptr || !bytes)
        return fail(HTB_ERR_INVALID, "null buffer");
    HTB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return HTB_OK;
}

int htb_host_unregister(void *ptr) {
    if (!ptr)
        return fail(HTB_ERR_INVALID, "null buffer");
    HTB_CUDA(cudaHostUnregister(ptr));
    return HTB_OK;
}

int htb_set_option(const char *key, int64_t value) {
    if (!key)
        return fail(HTB_ERR_INVALID, "null key");
    std::lock_guard<std::mutex> lock(g_option_mutex);
    auto it = g_options.find(key);
    if (it == g_options.end())
        return fail(HTB_ERR_INVALID, std::string("unknown option ") + key);
    it->second = value;
    return HTB_OK;
}

int htb_get_option(const char *key, int64_t *value) {
    if (!key || !value)
        return fail(HTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(g_option_mutex);
    auto it = g_options.find(key);
    if (it == g_options.end())
        return fail(HTB_ERR_INVALID, std::string("unknown option ") + key);
    *value = it->second;
    return HTB_OK;
}

// factors of the leaves compressed on the device (htb_create_compressed): where the packer's low-rank tasks find them
struct CompressedFactors {
    std::vector<AcaLeaf> leaves; // per leaf of the descriptor
    AcaPool pool;
};

static int create_impl(const htb_hmatrix_desc *desc, const htb_generator_desc *gen, const CompressedFactors *factors, htb_handle *out);

int htb_create(const htb_hmatrix_desc *desc, htb_handle *out) { return create_impl(desc, nullptr, nullptr, out); }

int htb_create_generated(const htb_hmatrix_desc *desc, const htb_generator_desc *gen, htb_handle *out) {
    if (!gen)
        return fail(HTB_ERR_INVALID, "null generator description");
    if (gen->spatial_dimension != 3)
        return fail(HTB_ERR_INVALID, "the built-in kernel functions are defined for points in R^3");
    if (gen->kernel < HTB_KERNEL_LAPLACE || gen->kernel > HTB_KERNEL_COMPLEX)
        return fail(HTB_ERR_INVALID, "unknown built-in kernel function");
    if (desc && kernel_is_complex(gen->kernel) != (desc->dtype == HTB_COMPLEX_DOUBLE))
        return fail(HTB_ERR_INVALID, "the kernel function does not produce the coefficient type of the H-matrix");
    if (desc && ((desc->nb_rows > 0 && !gen->target_points) || (desc->nb_cols > 0 && !gen->source_points)))
        return fail(HTB_ERR_INVALID, "null point array");
    return create_impl(desc, gen, nullptr, out);
}

int htb_download_store(htb_handle h, int side, void *dst, int64_t bytes) {
    if (!h || (side != 0 && side != 1) || !dst || bytes < 0)
        return fail(HTB_ERR_INVALID, "invalid argument");
    DeviceGuard guard(h->device);
    const int64_t have = static_cast<int64_t>(h->side_stream_bytes[side]);
    if (bytes > have)
        return fail(HTB_ERR_INVALID, "more bytes requested than the side's stream holds");
    if (bytes)
        HTB_CUDA(cudaMemcpy(dst, h->side[side].stream, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost));
    return HTB_OK;
}

// dense units of leaves that came without host data: generated on the device straight into the uploaded stream
static int generate_dense(htb_operator *h, const Packer &pk, const htb_generator_desc *gen) {
    const auto &tasks = pk.side[0].dense_tasks;
    if (tasks.empty())
        return HTB_OK;
    void *d_tasks = nullptr, *d_tp = nullptr, *d_sp = nullptr;
    auto cleanup = [&]() {
        for (void *p : {d_tasks, d_tp, d_sp})
            if (p)
                cudaFree(p);
    };
    const size_t tb = tasks.size() * sizeof(DenseTask), tpb = size_t(3) * h->nb_rows * sizeof(double), spb = size_t(3) * h->nb_cols * sizeof(double);
    cudaError_t e = cudaMalloc(&d_tasks, tb);
    if (e == cudaSuccess)
        e = cudaMalloc(&d_tp, std::max<size_t>(8, tpb));
    if (e == cudaSuccess)
        e = cudaMalloc(&d_sp, std::max<size_t>(8, spb));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_tasks, tasks.data(), tb, cudaMemcpyHostToDevice, h->own_stream);
    if (e == cudaSuccess && tpb)
        e = cudaMemcpyAsync(d_tp, gen->target_points, tpb, cudaMemcpyHostToDevice, h->own_stream);
    if (e == cudaSuccess && spb)
        e = cudaMemcpyAsync(d_sp, gen->source_points, spb, cudaMemcpyHostToDevice, h->own_stream);
    if (e == cudaSuccess)
        e = launch_generate_dense(gen->kernel, static_cast<const DenseTask *>(d_tasks), static_cast<long long>(tasks.size()), const_cast<unsigned char *>(h->side[0].stream), static_cast<const double *>(d_tp),
                                  static_cast<const double *>(d_sp), gen->wavenumber, h->own_stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(h->own_stream);
    cleanup();
    if (e != cudaSuccess)
        return cuda_fail(e, "dense leaf generation");
    h->launches++;
    h->generated_dense_units = static_cast<int64_t>(tasks.size());
    return HTB_OK;
}

// panels of the leaves compressed on the device: copied out of the factor pool into the uploaded streams of both sides
static int scatter_lowrank(htb_operator *h, const Packer &pk, const CompressedFactors &cf) {
    void *d_leaves = nullptr, *d_tasks = nullptr;
    cudaError_t e  = cudaMalloc(&d_leaves, std::max<size_t>(16, cf.leaves.size() * sizeof(AcaLeaf)));
    if (e == cudaSuccess && !cf.leaves.empty())
        e = cudaMemcpyAsync(d_leaves, cf.leaves.data(), cf.leaves.size() * sizeof(AcaLeaf), cudaMemcpyHostToDevice, h->own_stream);
    for (int s = 0; s < 2 && e == cudaSuccess; s++) {
        const auto &tasks = pk.side[s].lr_tasks;
        if (tasks.empty())
            continue;
        e = cudaMalloc(&d_tasks, tasks.size() * sizeof(DenseTask));
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * sizeof(DenseTask), cudaMemcpyHostToDevice, h->own_stream);
        if (e == cudaSuccess)
            e = launch_scatter_lowrank(h->dtype == HTB_COMPLEX_DOUBLE, static_cast<const DenseTask *>(d_tasks), static_cast<long long>(tasks.size()), s, const_cast<unsigned char *>(h->side[s].stream), static_cast<const AcaLeaf *>(d_leaves), cf.pool.pool,
                                       cf.pool.term_off, h->own_stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(h->own_stream);
        if (e == cudaSuccess)
            h->launches++;
        cudaFree(d_tasks);
        d_tasks = nullptr;
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(h->own_stream);
    cudaFree(d_leaves);
    if (e != cudaSuccess)
        return cuda_fail(e, "copy of the device-compressed factors into the leaf store");
    return HTB_OK;
}

static int create_impl(const htb_hmatrix_desc *desc, const htb_generator_desc *gen, const CompressedFactors *factors, htb_handle *out) {
    if (!desc || !out)
        return fail(HTB_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(HTB_ERR_CUDA, std::string("no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + "): htool_b200 has no CPU fallback");
    int device = desc->device;
    if (device < 0)
        HTB_CUDA(cudaGetDevice(&device));
    if (device >= ndev)
        return fail(HTB_ERR_INVALID, "device ordinal out of range");

    PackOptions popt;
    popt.generate_dense = gen != nullptr;
    popt.sort_units     = static_cast<int>(option("sort_units"));
    popt.block_rows  = static_cast<int>(option("block_rows"));
    popt.near_field  = option("m_near_field") != 0;
    popt.nf_rows     = static_cast<int>(std::max<int64_t>(0, std::min<int64_t>(128, option("m_nf_rows"))));
    popt.piece_cols  = static_cast<int>(option("piece_cols"));
    popt.stage_bytes = static_cast<int>(option("stage_bytes"));
    popt.cseg_bytes  = static_cast<int>(option("cseg_bytes"));
    popt.target_block_rows = static_cast<int>(option("target_block_rows"));
    popt.tail_split        = static_cast<int>(option("tail_split"));
    {
        cudaDeviceProp dp;
        if (cudaGetDeviceProperties(&dp, device) == cudaSuccess)
            popt.cta_slots = dp.multiProcessorCount * 3;
        if (option("cta_slots") > 0) // tests
            popt.cta_slots = static_cast<int>(option("cta_slots"));
    }
    std::unique_ptr<Packer> pk;
    const auto t_layout = std::chrono::steady_clock::now();
    try {
        pk = std::make_unique<Packer>(*desc, popt);
    } catch (const std::exception &ex) {
        return fail(HTB_ERR_INVALID, ex.what());
    }
    const double seconds_layout = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_layout).count();

    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(HTB_ERR_CUDA, "cudaSetDevice failed");
    auto h            = std::make_unique<htb_operator>();
    h->device         = device;
    h->dtype          = desc->dtype;
    h->esize          = pk->esize;
    h->nb_rows        = desc->nb_rows;
    h->nb_cols        = desc->nb_cols;
    h->row_offset     = desc->row_offset;
    h->col_offset     = desc->col_offset;
    h->symmetry       = desc->symmetry_for_leaves ? desc->symmetry_for_leaves : 'N';
    h->uplo           = desc->uplo_for_leaves ? desc->uplo_for_leaves : 'N';
    h->scratch_elems  = pk->scratch_elems;
    h->launch_cfg.block_rows  = pk->opt.block_rows; // resolved (0 = automatic)
    h->launch_cfg.stage_bytes = popt.stage_bytes;
    h->launch_cfg.cseg_bytes  = popt.cseg_bytes;
    h->launch_cfg.ring_stages        = static_cast<int>(option("ring_stages"));
    h->launch_cfg.reduce_ring_stages = static_cast<int>(option("reduce_ring_stages"));
    h->launch_cfg.evict_first = static_cast<int>(option("evict_first"));
    h->launch_cfg.reduce_blocks_per_cta = static_cast<int>(option("reduce_blocks_per_cta"));
    if (h->launch_cfg.ring_stages < 2 || h->launch_cfg.ring_stages > 32 || h->launch_cfg.reduce_ring_stages < 2 || h->launch_cfg.reduce_ring_stages > 32)
        return fail(HTB_ERR_INVALID, "ring_stages / reduce_ring_stages must be in [2, 32]");
    cudaDeviceProp prop{};
    HTB_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    if (std::max(apply_smem_bytes(h->launch_cfg, 16), reduce_smem_bytes(h->launch_cfg, 16)) > static_cast<size_t>(prop.sharedMemPerBlockOptin))
        return fail(HTB_ERR_INVALID, "the shared-memory ring (ring_stages x stage_bytes) exceeds the shared memory of an SM");
    h->launch_cfg.m_ring_stages        = static_cast<int>(option("m_ring_stages"));
    h->launch_cfg.m_reduce_ring_stages = static_cast<int>(option("m_reduce_ring_stages"));
    h->launch_cfg.m_b_ring_log2        = static_cast<int>(option("m_b_ring_log2"));
    h->launch_cfg.m_reduce_warps       = static_cast<int>(option("m_reduce_warps"));
    h->launch_cfg.m_b_producers        = static_cast<int>(option("m_b_producers"));
    if (h->launch_cfg.m_b_ring_log2 < 1 || h->launch_cfg.m_b_ring_log2 > 3)
        return fail(HTB_ERR_INVALID, "m_b_ring_log2 must be in [1, 3]");
    // a ring slot of the multi-RHS kernels = stage + the largest aux record of THIS store; ring depths 0 = as deep as the
    // shared memory of an SM allows (a slot is idle while it is refilled: bytes in flight are what keeps the DMMAs fed)
    h->launch_cfg.m_aux_bytes = static_cast<int>((std::max<uint32_t>(16u, std::max(pk->side[0].aux_max_bytes, pk->side[1].aux_max_bytes)) + 127u) & ~127u);
    h->launch_cfg.m_pad = static_cast<int>(std::min<int64_t>(8, std::max<int64_t>(0, option("m_pad")))) & ~1;
    for (int sd = 0; sd < 2; sd++)
        for (const BlockDesc &bd : pk->side[sd].blocks)
            h->launch_cfg.m_x_rows = std::max(h->launch_cfg.m_x_rows, static_cast<int>(bd.nrows));
    for (int *depth : {&h->launch_cfg.m_ring_stages, &h->launch_cfg.m_reduce_ring_stages}) {
        if (*depth != 0 && (*depth < 2 || *depth > 8))
            return fail(HTB_ERR_INVALID, "m_ring_stages / m_reduce_ring_stages must be 0 (automatic) or in [2, 8]");
        if (*depth == 0) {
            const bool red = depth == &h->launch_cfg.m_reduce_ring_stages;
            for (*depth = 8; *depth > 2; --*depth)
                if ((red ? reduce_m_smem_bytes(h->launch_cfg, 64, h->esize) : apply_m_smem_bytes(h->launch_cfg)) <= static_cast<size_t>(prop.sharedMemPerBlockOptin))
                    break;
        }
    }
    HTB_CUDA(configure_kernels(h->launch_cfg));
    h->m_path_ok = std::max(reduce_m_smem_bytes(h->launch_cfg, 64, h->esize), apply_m_smem_bytes(h->launch_cfg)) <= static_cast<size_t>(prop.sharedMemPerBlockOptin);
    if (h->m_path_ok)
        HTB_CUDA(configure_mkernels(h->launch_cfg, h->esize));
    h->mscratch_elems  = pk->mscratch_elems;
    g_pdl              = option("pdl") != 0;
    h->fused_symmetric = option("fused_symmetric") != 0 && fused_smem_bytes(h->launch_cfg, 16) <= static_cast<size_t>(prop.sharedMemPerBlockOptin);
    HTB_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;

    const auto t_upload = std::chrono::steady_clock::now();
    int rc              = upload_store(h.get(), *pk);
    const auto t_fill   = std::chrono::steady_clock::now();
    if (rc == HTB_OK && gen)
        rc = generate_dense(h.get(), *pk, gen);
    if (rc == HTB_OK && factors)
        rc = scatter_lowrank(h.get(), *pk, *factors);
    if (rc == HTB_OK && h->m_path_ok && !pk->nf.empty())
        h->nf_host = std::make_unique<NearFieldLayout>(std::move(pk->nf));
    h->create_seconds[0] = seconds_layout;
    h->create_seconds[1] = std::chrono::duration<double>(t_fill - t_upload).count();
    h->create_seconds[2] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_fill).count();
    if (rc != HTB_OK) {
        htb_destroy(h.release());
        return rc;
    }
    // scratch copies: one for the product, a second one for the transposed second application under symmetric storage
    const bool needs_second    = h->symmetry != 'N' && (h->side[0].any_twice || h->side[1].any_twice);
    h->needs_second_copy       = needs_second;
    const size_t scratch_bytes = std::max<size_t>(2, h->scratch_elems) * h->esize * (needs_second ? 2 : 1);
    e                          = cudaMalloc(&h->d_scratch, scratch_bytes);
    if (e != cudaSuccess) {
        htb_destroy(h.release());
        return cuda_fail(e, "cudaMalloc(scratch)");
    }
    cudaMemsetAsync(h->d_scratch, 0, scratch_bytes, h->own_stream);
    cudaStreamSynchronize(h->own_stream);
    h->workspace_bytes = scratch_bytes;

    htb_info &i             = h->info;
    i.nb_leaves             = pk->n_leaves;
    i.nb_dense_leaves       = pk->n_dense;
    i.nb_low_rank_leaves    = pk->n_lowrank;
    i.nb_leaves_applied_twice = pk->n_twice;
    i.coefficients          = pk->coefficients;
    i.coefficients_twice    = pk->coefficients_twice;
    i.rank_min              = pk->rank_min;
    i.rank_max              = pk->rank_max;
    i.dtype                 = h->dtype;
    i.device                = device;
    i.nb_rows               = h->nb_rows;
    i.nb_cols               = h->nb_cols;
    i.nb_target_blocks      = h->side[0].n_blocks;
    i.nb_source_blocks      = h->side[1].n_blocks;
    i.sm_count              = h->sm_count;
    *out                    = h.release();
    return HTB_OK;
}

// Leaf assembly on the device, second step: the admissible blocks are compressed by the batched ACA (aca.cu), the ranks come
// back to the host (they decide the layout of the store), the packer lays the store out with empty low-rank panels and the
// factors are copied from the pool into the uploaded streams. The factors never exist on the host.
int htb_create_compressed(const htb_hmatrix_desc *desc, const htb_generator_desc *gen, double epsilon, htb_handle *out) {
    const auto t_begin = std::chrono::steady_clock::now();
    if (!desc || !gen || !out)
        return fail(HTB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (gen->spatial_dimension != 3)
        return fail(HTB_ERR_INVALID, "the built-in kernel functions are defined for points in R^3");
    if (gen->kernel < HTB_KERNEL_LAPLACE || gen->kernel > HTB_KERNEL_COMPLEX)
        return fail(HTB_ERR_INVALID, "unknown built-in kernel function");
    if (kernel_is_complex(gen->kernel) != (desc->dtype == HTB_COMPLEX_DOUBLE))
        return fail(HTB_ERR_INVALID, "the kernel function does not produce the coefficient type of the H-matrix");
    if (!(epsilon > 0.))
        return fail(HTB_ERR_INVALID, "epsilon must be positive (a required rank is not supported on the device)");
    if ((desc->nb_rows > 0 && !gen->target_points) || (desc->nb_cols > 0 && !gen->source_points))
        return fail(HTB_ERR_INVALID, "null point array");
    if (desc->nb_leaves < 0 || (desc->nb_leaves > 0 && !desc->leaves))
        return fail(HTB_ERR_INVALID, "null leaf array");

    // the blocks to compress, one size class after the other (teams of 512 / 128 / 32 threads), large blocks first
    std::vector<AcaBlock> blocks;
    std::vector<htb_leaf> leaves(desc->leaves, desc->leaves + desc->nb_leaves);
    uint64_t term_slots = 0;
    for (int64_t i = 0; i < desc->nb_leaves; i++) {
        const htb_leaf &l = leaves[i];
        if (l.rank != HTB_RANK_COMPRESS)
            continue;
        if (l.data0 || l.data1)
            return fail(HTB_ERR_INVALID, "a leaf to compress carries data");
        if (l.nb_rows < 0 || l.nb_cols < 0 || l.row_offset < 0 || l.col_offset < 0 || int64_t(l.row_offset) + l.nb_rows > desc->nb_rows || int64_t(l.col_offset) + l.nb_cols > desc->nb_cols)
            return fail(HTB_ERR_INVALID, "leaf outside the root block");
        if (l.nb_rows == 0 || l.nb_cols == 0) {
            leaves[i].rank = 0;
            continue;
        }
        AcaBlock b{};
        b.lrow = l.row_offset, b.lcol = l.col_offset, b.m = l.nb_rows, b.n = l.nb_cols;
        b.swapped  = (int64_t(desc->row_offset) + l.row_offset >= int64_t(desc->col_offset) + l.col_offset) ? 0 : 1; // sympartialACA.hpp:46
        b.term_cap = static_cast<uint16_t>(std::min<int64_t>(aca_max_rank(aca_team(b.m, b.n)), (int64_t(b.m) * b.n) / (int64_t(b.m) + b.n)));
        b.leaf     = static_cast<uint32_t>(i);
        blocks.push_back(b);
    }
    if (blocks.empty()) {
        htb_hmatrix_desc d2 = *desc; // nothing to compress
        d2.leaves           = leaves.data();
        int rc = create_impl(&d2, gen, nullptr, out);
        if (rc == HTB_OK) {
            (*out)->leaf_ranks.resize(leaves.size());
            for (size_t i = 0; i < leaves.size(); i++)
                (*out)->leaf_ranks[i] = leaves[i].rank;
            (*out)->compression.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        }
        return rc;
    }
    auto team_of = [](const AcaBlock &b) { return aca_team(b.m, b.n); };
    std::stable_sort(blocks.begin(), blocks.end(), [&](const AcaBlock &a, const AcaBlock &b) {
        const int ta = team_of(a), tb = team_of(b);
        return ta != tb ? ta > tb : int64_t(a.m) + a.n > int64_t(b.m) + b.n;
    });
    for (AcaBlock &b : blocks) {
        b.term_base = static_cast<uint32_t>(term_slots);
        term_slots += b.term_cap;
    }
    if (term_slots >= (uint64_t(1) << 32))
        return fail(HTB_ERR_UNSUPPORTED, "term table of the device compression exceeds 2^32 slots");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(HTB_ERR_CUDA, std::string("no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + "): htool_b200 has no CPU fallback");
    int device = desc->device;
    if (device < 0)
        HTB_CUDA(cudaGetDevice(&device));
    if (device >= ndev)
        return fail(HTB_ERR_INVALID, "device ordinal out of range");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(HTB_ERR_CUDA, "cudaSetDevice failed");

    CompressedFactors cf;
    void *d_blocks = nullptr, *d_tp = nullptr, *d_sp = nullptr, *d_rank = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_class[3] = {nullptr, nullptr, nullptr};
    int class_slot[3]       = {0, 0, 0};
    double seconds_team[3]  = {0., 0., 0.};
    int64_t blocks_team[3]  = {0, 0, 0};
    auto cleanup = [&]() {
        for (void *p : {d_blocks, d_tp, d_sp, d_rank, static_cast<void *>(cf.pool.pool), static_cast<void *>(cf.pool.cursor), static_cast<void *>(cf.pool.term_off)})
            if (p)
                cudaFree(p);
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
        for (cudaEvent_t ev : ev_class)
            if (ev)
                cudaEventDestroy(ev);
        if (st)
            cudaStreamDestroy(st);
    };
#define HTB_CC(call)                          \
    do {                                      \
        cudaError_t e__ = (call);             \
        if (e__ != cudaSuccess) {             \
            cleanup();                        \
            return cuda_fail(e__, #call);     \
        }                                     \
    } while (0)
    const size_t tpb = size_t(3) * desc->nb_rows * sizeof(double), spb = size_t(3) * desc->nb_cols * sizeof(double);
    HTB_CC(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    HTB_CC(cudaEventCreate(&ev0));
    HTB_CC(cudaEventCreate(&ev1));
    for (cudaEvent_t &ev : ev_class)
        HTB_CC(cudaEventCreate(&ev));
    HTB_CC(cudaMalloc(&d_blocks, blocks.size() * sizeof(AcaBlock)));
    HTB_CC(cudaMalloc(&d_rank, blocks.size() * sizeof(int32_t)));
    HTB_CC(cudaMalloc(&d_tp, std::max<size_t>(8, tpb)));
    HTB_CC(cudaMalloc(&d_sp, std::max<size_t>(8, spb)));
    HTB_CC(cudaMalloc(reinterpret_cast<void **>(&cf.pool.cursor), sizeof(unsigned long long)));
    HTB_CC(cudaMalloc(reinterpret_cast<void **>(&cf.pool.term_off), std::max<uint64_t>(1, term_slots) * sizeof(uint32_t)));
    HTB_CC(cudaMemcpyAsync(d_blocks, blocks.data(), blocks.size() * sizeof(AcaBlock), cudaMemcpyHostToDevice, st));
    HTB_CC(cudaMemcpyAsync(d_tp, gen->target_points, tpb, cudaMemcpyHostToDevice, st));
    HTB_CC(cudaMemcpyAsync(d_sp, gen->source_points, spb, cudaMemcpyHostToDevice, st));

    const auto t_prepared = std::chrono::steady_clock::now();
    // The factor pool. The ranks are unknown before the compression: the pool is sized for `guess` terms per block (never
    // more than a block can hold) and the whole batch is repeated with twice the guess when it overflows.
    size_t free_bytes = 0, total_bytes = 0;
    HTB_CC(cudaMemGetInfo(&free_bytes, &total_bytes));
    const uint64_t pool_limit = std::min<uint64_t>(uint64_t(1) << 33, static_cast<uint64_t>(free_bytes * 0.45) / sizeof(double)); // (term offsets are stored in 16-byte units, 32 bits)
    std::vector<int32_t> rank(blocks.size());
    double seconds_aca = 0.;
    const int fma_axpy  = option("aca_fma_axpy") != 0;
    const int dots_mode = static_cast<int>(option("aca_dots"));
    for (int64_t guess = std::max<int64_t>(1, option("aca_rank_guess"));; guess *= 2) {
        uint64_t want = 0, full = 0;
        for (const AcaBlock &b : blocks) {
            const uint64_t len = desc->dtype == HTB_COMPLEX_DOUBLE ? 2u * (uint64_t(b.m) + b.n) : (uint64_t(b.m) + b.n + 1u) & ~uint64_t(1); // doubles per term
            want += len * std::min<uint64_t>(b.term_cap, static_cast<uint64_t>(guess));
            full += len * b.term_cap;
        }
        const uint64_t capacity = std::max<uint64_t>(2, std::min(want, pool_limit));
        if (cf.pool.pool)
            cudaFree(cf.pool.pool);
        cf.pool.pool = nullptr;
        HTB_CC(cudaMalloc(reinterpret_cast<void **>(&cf.pool.pool), capacity * sizeof(double)));
        cf.pool.capacity = capacity;
        HTB_CC(cudaMemsetAsync(cf.pool.cursor, 0, sizeof(unsigned long long), st));
        HTB_CC(cudaEventRecord(ev0, st));
        int n_class = 0;
        for (size_t first = 0; first < blocks.size();) { // one launch per size class
            size_t last = first;
            while (last < blocks.size() && team_of(blocks[last]) == team_of(blocks[first]))
                last++;
            HTB_CC(launch_aca(gen->kernel, team_of(blocks[first]), static_cast<const AcaBlock *>(d_blocks), static_cast<long long>(first), static_cast<long long>(last - first), static_cast<const double *>(d_tp),
                              static_cast<const double *>(d_sp), gen->wavenumber, epsilon, fma_axpy, cf.pool, static_cast<int32_t *>(d_rank), dots_mode, st));
            const int slot = team_of(blocks[first]) == 512 ? 0 : (team_of(blocks[first]) == 128 ? 1 : 2);
            HTB_CC(cudaEventRecord(ev_class[n_class], st));
            class_slot[n_class++] = slot;
            blocks_team[slot]     = static_cast<int64_t>(last - first);
            first = last;
        }
        HTB_CC(cudaEventRecord(ev1, st));
        HTB_CC(cudaMemcpyAsync(rank.data(), d_rank, rank.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        HTB_CC(cudaStreamSynchronize(st));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        seconds_aca += ms * 1e-3;
        for (int c = 0; c < n_class; c++) {
            cudaEventElapsedTime(&ms, c == 0 ? ev0 : ev_class[c - 1], ev_class[c]);
            seconds_team[class_slot[c]] = ms * 1e-3; // (of the last attempt)
        }
        bool overflow = false;
        for (int32_t q : rank)
            overflow = overflow || q == kAcaPoolOverflow;
        if (!overflow)
            break;
        if (capacity >= std::min(full, pool_limit)) {
            cleanup();
            return fail(HTB_ERR_CUDA, "the factor pool of the device compression does not fit the device memory");
        }
    }

    const auto t_compressed = std::chrono::steady_clock::now();
    htb_compression_info ci{};
    ci.seconds_prepare  = std::chrono::duration<double>(t_prepared - t_begin).count();
    ci.seconds_compress = std::chrono::duration<double>(t_compressed - t_prepared).count();
    ci.nb_blocks  = static_cast<int64_t>(blocks.size());
    ci.rank_min   = INT32_MAX;
    ci.pool_bytes = static_cast<int64_t>(cf.pool.capacity * sizeof(double));
    cf.leaves.assign(leaves.size(), AcaLeaf{0u, 0u, 0u, 0u});
    for (size_t i = 0; i < blocks.size(); i++) {
        const AcaBlock &b = blocks[i];
        const int32_t q   = rank[i];
        if (q == kAcaRankCap) {
            cleanup();
            return fail(HTB_ERR_UNSUPPORTED, "an admissible block needs more than " + std::to_string(aca_max_rank(512)) + " terms at this epsilon: compress on the host");
        }
        if (q > 0) {
            leaves[b.leaf].rank = q;
            cf.leaves[b.leaf]   = AcaLeaf{b.term_base, static_cast<uint32_t>(b.swapped ? b.n : b.m), b.swapped, 0u};
            ci.coefficients += int64_t(q) * (int64_t(b.m) + b.n);
            ci.rank_min = std::min(ci.rank_min, q);
            ci.rank_max = std::max(ci.rank_max, q);
        } else { // the compression failed: a dense leaf, generated like the others (tree_builder.hpp:619-625)
            leaves[b.leaf].rank = -1;
            ci.nb_failed++;
        }
    }
    if (ci.rank_min == INT32_MAX)
        ci.rank_min = 0;
    ci.seconds_aca = seconds_aca;
    for (void **p : {&d_blocks, &d_tp, &d_sp, &d_rank}) { // (the store needs the room)
        cudaFree(*p);
        *p = nullptr;
    }

    htb_hmatrix_desc d2 = *desc;
    d2.leaves           = leaves.data();
    d2.device           = device;
    const auto t_store  = std::chrono::steady_clock::now();
    int rc              = create_impl(&d2, gen, &cf, out);
    ci.seconds_store    = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_store).count();
    cleanup();
#undef HTB_CC
    if (rc != HTB_OK)
        return rc;
    (*out)->leaf_ranks.resize(leaves.size());
    for (size_t i = 0; i < leaves.size(); i++)
        (*out)->leaf_ranks[i] = leaves[i].rank;
    ci.seconds_total      = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    for (int c = 0; c < 3; c++) {
        ci.seconds_aca_team[c] = seconds_team[c];
        ci.nb_blocks_team[c]   = blocks_team[c];
    }
    (*out)->compression   = ci;
    return HTB_OK;
}

int htb_get_leaf_ranks(htb_handle h, int32_t *ranks, int64_t nb_leaves) {
    if (!h || !ranks)
        return fail(HTB_ERR_INVALID, "null argument");
    if (h->leaf_ranks.empty() && nb_leaves != 0)
        return fail(HTB_ERR_INVALID, "the operator was not created by htb_create_compressed");
    if (nb_leaves != static_cast<int64_t>(h->leaf_ranks.size()))
        return fail(HTB_ERR_INVALID, "leaf count differs from the descriptor's");
    std::copy(h->leaf_ranks.begin(), h->leaf_ranks.end(), ranks);
    return HTB_OK;
}

int htb_get_compression_info(htb_handle h, htb_compression_info *info) {
    if (!h || !info)
        return fail(HTB_ERR_INVALID, "null argument");
    *info                = h->compression;
    info->seconds_layout = h->create_seconds[0];
    info->seconds_upload = h->create_seconds[1];
    info->seconds_fill   = h->create_seconds[2];
    return HTB_OK;
}

int htb_destroy(htb_handle h) {
    if (!h)
        return HTB_OK;
    DeviceGuard guard(h->device);
    dist_destroy(h);
    if (h->own_stream)
        cudaStreamSynchronize(h->own_stream);
    for (void *p : h->owned)
        cudaFree(p);
    for (void *p : {h->d_mscratch, h->d_mstage, h->d_scratch, h->d_in, h->d_out, h->d_perm[0], h->d_perm[1], h->d_work_in, h->d_work_out, h->d_krylov})
        if (p)
            cudaFree(p);
    for (void *p : {h->h_in, h->h_out, h->h_krylov})
        if (p)
            cudaFreeHost(p);
    for (cudaEvent_t ev : h->chunk_events)
        cudaEventDestroy(ev);
    if (h->own_stream)
        cudaStreamDestroy(h->own_stream);
    delete h;
    return HTB_OK;
}

int htb_get_info(htb_handle h, htb_info *info) {
    if (!h || !info)
        return fail(HTB_ERR_INVALID, "null argument");
    *info                  = h->info;
    info->dist_gather      = dist_gather_mode(h);
    info->store_bytes      = static_cast<int64_t>(h->store_bytes);
    info->descriptor_bytes = static_cast<int64_t>(h->descriptor_bytes);
    info->workspace_bytes  = static_cast<int64_t>(h->workspace_bytes + h->in_cap + h->out_cap + h->work_cap * 2 + h->krylov_cap);
    return HTB_OK;
}

int htb_set_stream(htb_handle h, void *cuda_stream) {
    if (!h)
        return fail(HTB_ERR_INVALID, "null handle");
    h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return HTB_OK;
}

int htb_synchronize(htb_handle h) {
    if (!h)
        return fail(HTB_ERR_INVALID, "null handle");
    DeviceGuard guard(h->device);
    HTB_CUDA(cudaStreamSynchronize(h->stream));
    return HTB_OK;
}

int htb_launch_count(htb_handle h, int64_t *count) {
    if (!h || !count)
        return fail(HTB_ERR_INVALID, "null argument");
    *count = h->launches;
    return HTB_OK;
}

int htb_profile_passes(htb_handle h, int enable) {
    if (!h)
        return fail(HTB_ERR_INVALID, "null handle");
    h->profiling = enable != 0;
    return HTB_OK;
}

int htb_get_pass_times(htb_handle h, double ms[HTB_PASS_KINDS], int64_t launches[HTB_PASS_KINDS]) {
    if (!h || !ms || !launches)
        return fail(HTB_ERR_INVALID, "null argument");
    DeviceGuard guard(h->device);
    HTB_CUDA(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < HTB_PASS_KINDS; k++) {
        ms[k]       = 0;
        launches[k] = 0;
    }
    for (auto &tl : h->timed) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, tl.start, tl.stop) == cudaSuccess) {
            ms[tl.kind] += t;
            launches[tl.kind]++;
        }
        cudaEventDestroy(tl.start);
        cudaEventDestroy(tl.stop);
    }
    h->timed.clear();
    return HTB_OK;
}

int htb_add_vector_product(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mem_kind) {
    return htb_add_matrix_product_row_major(h, trans, alpha, in, beta, out, 1, mem_kind);
}

int htb_add_matrix_product_row_major(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, int mem_kind) {
    if (!h || !alpha || !beta || mu < 0 || (mem_kind != HTB_MEM_HOST && mem_kind != HTB_MEM_DEVICE))
        return fail(HTB_ERR_INVALID, "invalid argument");
    if (mu == 0)
        return HTB_OK;
    if (!in || !out)
        return fail(HTB_ERR_INVALID, "null vector");
    DeviceGuard guard(h->device);
    if (mem_kind == HTB_MEM_DEVICE)
        return product_device(h, trans, alpha, in, beta, out, mu);
    return product_host(h, trans, alpha, in, beta, out, mu);
}

int htb_set_permutations(htb_handle h, const int32_t *target_permutation, const int32_t *source_permutation) {
    if (!h || (!target_permutation && h->nb_rows > 0) || (!source_permutation && h->nb_cols > 0)) // (an empty side has nothing to permute)
        return fail(HTB_ERR_INVALID, "null argument");
    DeviceGuard guard(h->device);
    const int32_t *src[2] = {target_permutation, source_permutation};
    const int n[2]        = {h->nb_rows, h->nb_cols};
    for (int s = 0; s < 2; s++) {
        for (int i = 0; i < n[s]; i++)
            if (src[s][i] < 0 || src[s][i] >= n[s])
                return fail(HTB_ERR_INVALID, "permutation entry out of range");
        if (h->d_perm[s])
            cudaFree(h->d_perm[s]);
        h->d_perm[s] = nullptr;
        HTB_CUDA(cudaMalloc(&h->d_perm[s], std::max<size_t>(1, n[s]) * sizeof(int32_t)));
        if (n[s] > 0)
            HTB_CUDA(cudaMemcpy(h->d_perm[s], src[s], n[s] * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    return HTB_OK;
}

// add_hmatrix_vector_product / add_hmatrix_matrix_product in user numbering: gather in (and out when beta != 0)
// into cluster numbering, product, scatter out (add_hmatrix_vector_product.hpp:173-197, add_hmatrix_matrix_product.hpp:176-205)
static int product_user(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, int mem_kind, bool colmajor) {
    if (!h || !alpha || !beta || !in || !out || mu < 1)
        return fail(HTB_ERR_INVALID, "invalid argument");
    if (!h->d_perm[0] || !h->d_perm[1])
        return fail(HTB_ERR_INVALID, "htb_set_permutations has not been called");
    int rc = check_trans(h, trans);
    if (rc != HTB_OK)
        return rc;
    DeviceGuard guard(h->device);
    const int ni = trans == 'N' ? h->nb_cols : h->nb_rows, no = trans == 'N' ? h->nb_rows : h->nb_cols;
    const int32_t *pin  = static_cast<const int32_t *>(trans == 'N' ? h->d_perm[1] : h->d_perm[0]);
    const int32_t *pout = static_cast<const int32_t *>(trans == 'N' ? h->d_perm[0] : h->d_perm[1]);
    const size_t in_bytes = size_t(ni) * mu * h->esize, out_bytes = size_t(no) * mu * h->esize;
    const size_t need = std::max(in_bytes, out_bytes);
    if (need > h->work_cap) {
        for (void **p : {&h->d_work_in, &h->d_work_out}) {
            if (*p)
                cudaFree(*p);
            *p = nullptr;
        }
        h->work_cap = 0;
        HTB_CUDA(cudaMalloc(&h->d_work_in, need));
        HTB_CUDA(cudaMalloc(&h->d_work_out, need));
        h->work_cap = need;
    }
    cudaStream_t st = h->stream;
    const void *din = in;
    void *dout      = out;
    if (mem_kind == HTB_MEM_HOST) {
        if ((rc = ensure_staging(h, in_bytes, out_bytes)) != HTB_OK)
            return rc;
        if ((rc = staged_h2d(h, h->d_in, h->h_in, in, in_bytes, st)) != HTB_OK)
            return rc;
        if (!beta_is_zero(h, beta) && (rc = staged_h2d(h, h->d_out, h->h_out, out, out_bytes, st)) != HTB_OK)
            return rc;
        din  = h->d_in;
        dout = h->d_out;
    }
    const bool b0 = beta_is_zero(h, beta);
    if (h->dtype == HTB_DOUBLE) {
        HTB_CUDA(launch_permute<double>(static_cast<const double *>(din), static_cast<double *>(h->d_work_in), pin, ni, mu, true, colmajor, st));
        if (!b0)
            HTB_CUDA(launch_permute<double>(static_cast<const double *>(dout), static_cast<double *>(h->d_work_out), pout, no, mu, true, colmajor, st));
    } else {
        HTB_CUDA(launch_permute<cplx>(static_cast<const cplx *>(din), static_cast<cplx *>(h->d_work_in), pin, ni, mu, true, colmajor, st));
        if (!b0)
            HTB_CUDA(launch_permute<cplx>(static_cast<const cplx *>(dout), static_cast<cplx *>(h->d_work_out), pout, no, mu, true, colmajor, st));
    }
    h->launches += b0 ? 1 : 2;
    if ((rc = product_device(h, trans, alpha, h->d_work_in, beta, h->d_work_out, mu)) != HTB_OK)
        return rc;
    if (h->dtype == HTB_DOUBLE)
        HTB_CUDA(launch_permute<double>(static_cast<const double *>(h->d_work_out), static_cast<double *>(dout), pout, no, mu, false, colmajor, st));
    else
        HTB_CUDA(launch_permute<cplx>(static_cast<const cplx *>(h->d_work_out), static_cast<cplx *>(dout), pout, no, mu, false, colmajor, st));
    h->launches++;
    if (mem_kind == HTB_MEM_HOST)
        return staged_d2h(h, out, h->h_out, h->d_out, out_bytes, st);
    return HTB_OK;
}

struct PackedOwner {
    SideLayout layout;
    std::vector<char> stream, headers;
    std::vector<uint64_t> header_offsets;
};

int htb_pack_host(const htb_hmatrix_desc *desc, int side, htb_packed_side *out) {
    if (!desc || !out || (side != 0 && side != 1))
        return fail(HTB_ERR_INVALID, "invalid argument");
    PackOptions popt;
    popt.block_rows  = static_cast<int>(option("block_rows"));
    popt.near_field  = option("m_near_field") != 0;
    popt.nf_rows     = static_cast<int>(std::max<int64_t>(0, std::min<int64_t>(128, option("m_nf_rows"))));
    popt.piece_cols  = static_cast<int>(option("piece_cols"));
    popt.stage_bytes = static_cast<int>(option("stage_bytes"));
    popt.cseg_bytes  = static_cast<int>(option("cseg_bytes"));
    popt.target_block_rows = static_cast<int>(option("target_block_rows"));
    popt.tail_split        = static_cast<int>(option("tail_split"));
    if (option("cta_slots") > 0)
        popt.cta_slots = static_cast<int>(option("cta_slots"));
    popt.generate_dense = option("pack_generate_dense") != 0; // tests: what htb_create_generated would upload
    popt.sort_units     = static_cast<int>(option("sort_units"));
    try {
        Packer pk(*desc, popt);
        auto *own   = new PackedOwner();
        own->layout = pk.side[side];
        own->stream.resize(pk.side[side].stream_bytes);
        if (!own->layout.blocks.empty()) {
            const auto t0 = std::chrono::steady_clock::now();
            pk.fill(side, 0, static_cast<int>(own->layout.blocks.size()), own->stream.data());
            if (std::getenv("HTB_PACK_TIMING")) // development aid (tools/pack_time.py --with-data)
                std::fprintf(stderr, "[htb pack] fill side %d: %.3f s for %.2f GB\n", side, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), pk.side[side].stream_bytes / 1e9);
        }
        out->n             = own->layout.n;
        out->n_blocks      = static_cast<int32_t>(own->layout.blocks.size());
        out->n_stages      = static_cast<int64_t>(own->layout.stages.size());
        out->n_combine     = static_cast<int64_t>(own->layout.combine.size());
        out->n_combine_dst = static_cast<int64_t>(own->layout.combine_dst.size());
        out->stream_bytes  = static_cast<int64_t>(own->stream.size());
        out->scratch_elems = static_cast<int64_t>(pk.scratch_elems);
        out->cs_base       = static_cast<int64_t>(own->layout.cs_base);
        out->cs_elems      = static_cast<int64_t>(own->layout.cs_elems);
        out->part_base     = static_cast<int64_t>(own->layout.part_base);
        out->part_elems    = static_cast<int64_t>(own->layout.part_elems);
        out->piece_cols    = pk.piece;
        out->block_rows    = pk.opt.block_rows;
        out->n_munits      = static_cast<int64_t>(own->layout.munits.size());
        out->n_combine_m   = static_cast<int64_t>(own->layout.combine_m.size());
        out->mscratch_elems = static_cast<int64_t>(pk.mscratch_elems);
        out->munits        = own->layout.munits.data();
        out->combine_m     = own->layout.combine_m.data();
        out->combine_dst   = own->layout.combine_dst.data();
        out->blocks        = own->layout.blocks.data();
        out->stages        = own->layout.stages.data();
        out->order         = own->layout.order.data();
        out->combine       = own->layout.combine.data();
        out->stream        = own->stream.data();
        out->owner         = own;
        out->aux_bytes     = static_cast<int64_t>(own->layout.aux_reduce.size());
        out->aux_reduce    = own->layout.aux_reduce.data();
        out->aux_apply     = own->layout.aux_apply.data();
        out->n_dense_tasks = static_cast<int64_t>(own->layout.dense_tasks.size());
        out->dense_tasks   = own->layout.dense_tasks.data();
        out->n_lowrank_tasks = static_cast<int64_t>(own->layout.lr_tasks.size());
        out->lowrank_tasks   = own->layout.lr_tasks.data();
        out->header_bytes = 0, out->headers = nullptr, out->header_offsets = nullptr;
        if (pk.all_on_device && !own->layout.blocks.empty()) { // what upload_store sends instead of the stream (upload_headers_only)
            own->header_offsets = pk.header_offsets(side);
            own->headers.resize(own->header_offsets.back());
            pk.fill_headers(side, 0, static_cast<int>(own->layout.blocks.size()), own->headers.data());
            out->header_bytes   = static_cast<int64_t>(own->headers.size());
            out->headers        = own->headers.data();
            out->header_offsets = own->header_offsets.data();
        }
    } catch (const std::exception &ex) {
        return fail(HTB_ERR_INVALID, ex.what());
    }
    return HTB_OK;
}

int htb_pack_free(htb_packed_side *packed) {
    if (packed && packed->owner) {
        delete static_cast<PackedOwner *>(packed->owner);
        packed->owner = nullptr;
    }
    return HTB_OK;
}

int htb_add_vector_product_user_numbering(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mem_kind) {
    return product_user(h, trans, alpha, in, beta, out, 1, mem_kind, false);
}

int htb_add_matrix_product_user_numbering(htb_handle h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu, int mem_kind) {
    return product_user(h, trans, alpha, in, beta, out, mu, mem_kind, true);
}

} // extern "C"

// tools/dmma_probe.cu — how many warps / independent accumulators keep the FP64 tensor pipe of one SM busy when the
// fragments come from shared memory (the situation of REDUCE_M / APPLY_M), and what warps that spin on an mbarrier cost
// the working ones. Development probe, prints one JSON line per configuration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// work warps: per k-step one B fragment + ILP A fragments from shared memory, ILP DMMAs on ILP accumulators.
// PIPE: fragments of the next k-step loaded before the DMMAs of this one. spin warps poll an mbarrier until the workers are done.
template <int ILP, bool PIPE>
__global__ void __launch_bounds__(1024) probe(double *out, int ksteps, int work_warps) {
    extern __shared__ double sm[];
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x)
        sm[i] = 1.0 + 1e-9 * i;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(work_warps) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= work_warps) { // spinning warp
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u), "r"(0x989680u) : "memory");
        } while (!done);
        return;
    }
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++)
        c[i][0] = c[i][1] = 0.;
    const double *xa = sm + (lane & 3) * 68 + (lane >> 2) + warp * 8;
    const double *pb = sm + 4096 + (lane >> 2) * 124 + (lane & 3);
    if (PIPE) {
        double a[ILP], b = pb[0];
#pragma unroll
        for (int i = 0; i < ILP; i++)
            a[i] = xa[8 * i];
#pragma unroll 2
        for (int k = 0; k < ksteps; k++) {
            const int off = ((k + 1) & 7) * 4;
            double an[ILP], bn = pb[off];
#pragma unroll
            for (int i = 0; i < ILP; i++)
                an[i] = xa[8 * i + off * 68];
#pragma unroll
            for (int i = 0; i < ILP; i++)
                dmma(c[i], a[i], b);
            b = bn;
#pragma unroll
            for (int i = 0; i < ILP; i++)
                a[i] = an[i];
        }
    } else {
        for (int k = 0; k < ksteps; k++) {
            const int off  = (k & 7) * 4;
            const double b = pb[off];
#pragma unroll
            for (int i = 0; i < ILP; i++)
                dmma(c[i], xa[8 * i + off * 68], b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
        s += c[i][0] + c[i][1];
    out[(blockIdx.x * blockDim.x + threadIdx.x)] = s;
    __syncwarp();
    if (lane == 0)
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
}

template <int ILP, bool PIPE>
static void run(double *out, int sms, int work, int spin) {
    const int ksteps = 40000 / ILP * 8 / 8;
    const int threads = (work + spin) * 32;
    cudaFuncSetAttribute(probe<ILP, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<ILP, PIPE><<<sms, threads, 65536>>>(out, ksteps, work);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<ILP, PIPE><<<sms, threads, 65536>>>(out, ksteps, work);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = double(sms) * work * ksteps * ILP * 512.0 / (ms * 1e-3) / 1e12;
    printf("{\"ilp\": %d, \"pipelined\": %d, \"work_warps\": %d, \"spin_warps\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", ILP, PIPE ? 1 : 0, work, spin, ms, tf, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess)
        return 1;
    double *out;
    cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 1024);
    const int sms = prop.multiProcessorCount;
    for (int work : {1, 2, 4, 8, 12, 16, 24}) {
        run<8, false>(out, sms, work, 0);
        run<8, true>(out, sms, work, 0);
        run<4, true>(out, sms, work, 0);
        run<2, true>(out, sms, work, 0);
    }
    for (int spin : {4, 8, 12, 16}) {
        run<8, true>(out, sms, 4, spin);
        run<8, true>(out, sms, 8, spin);
        run<8, false>(out, sms, 8, spin);
    }
    cudaFree(out);
    return 0;
}

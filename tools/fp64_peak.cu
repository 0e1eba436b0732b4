// tools/fp64_peak.cu — measures the FP64 peaks of the GPU it runs on: DMMA (mma.sync m8n8k4 f64, the only FP64
// tensor-core shape on sm_100a: SURVEY.md 7 "FP64 tensor path") and plain DFMA. MEASURED_PEAKS.json has no FP64
// figure, and the multi-RHS roofline (BASELINE.json config 3) needs one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
// Prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double *out, int iters, double a0, double b0) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++)
        c[i][0] = c[i][1] = 0.;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
        s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a0, double b0) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++)
        c[i] = i;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
        s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F &&launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    const int sms = prop.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    constexpr int ILP = 8;
    const double ms_mma = time_ms([&] { dmma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    const double ms_fma = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 0.9999999, 1e-3); });
    // one m8n8k4 DMMA = 8*8*4 FMA = 512 flop per warp; one DFMA = 2 flop per thread
    const double warps = double(blocks) * threads / 32;
    const double tf_mma = warps * iters * ILP * 512.0 / (ms_mma * 1e-3) / 1e12;
    const double tf_fma = double(blocks) * threads * iters * ILP * 2.0 / (ms_fma * 1e-3) / 1e12;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_tflops\": %.2f, \"dmma_ms\": %.3f, \"dfma_ms\": %.3f}\n", prop.name, sms, tf_mma, tf_fma, ms_mma, ms_fma);
    cudaFree(out);
    return 0;
}

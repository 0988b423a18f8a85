// fp32_peak.cu -- the FP32 (non-tensor) FMA peak of this GPU, which
// MEASURED_PEAKS.json does not hold (SURVEY 8d asks for it next to the HBM
// number: roofline time = max(bytes / HBM_BW, FMAs / FP32_peak)).  Two kernels of
// independent register-resident FMA chains: scalar `fma.rn.f32` (FFMA) and the
// packed `fma.rn.f32x2` (FFMA2) the resize kernels issue.  CUDA events, best of
// 5 launches after a warm-up.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp32_peak tools/fp32_peak.cu
#include <cuda_runtime.h>

#include <cstdio>

constexpr int CHAINS = 16;   // independent accumulators per thread
constexpr int ITERS  = 4096; // FMAs per chain

__global__ void __launch_bounds__(256)
ffma_kernel(float* out, float a, float b)
{
    float acc[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k)
        acc[k] = float(threadIdx.x + k);
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k)
            acc[k] = __fmaf_rn(acc[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k)
        s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256)
ffma2_kernel(float* out, float a, float b)
{
    unsigned long long acc[CHAINS];
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        float x = float(threadIdx.x + k);
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(x));
    }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(aa), "l"(bb));
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<class K>
static double
best_ms(K kernel, int blocks, float* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<blocks, 256>>>(out, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        kernel<<<blocks, 256>>>(out, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best)
            best = ms;
    }
    return best;
}

int
main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int blocks = p.multiProcessorCount * 8;
    float* out       = nullptr;
    cudaMalloc(&out, size_t(blocks) * 256 * sizeof(float));
    const double threads = double(blocks) * 256;
    const double ms1 = best_ms(ffma_kernel, blocks, out);
    const double ms2 = best_ms(ffma2_kernel, blocks, out);
    const double f1  = threads * CHAINS * ITERS * 2.0 / (ms1 * 1e-3) / 1e12;
    const double f2  = threads * CHAINS * ITERS * 4.0 / (ms2 * 1e-3) / 1e12;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %d, "
           "\"ffma_tflops\": %.2f, \"ffma_ms\": %.4f, \"ffma2_tflops\": %.2f, \"ffma2_ms\": %.4f, "
           "\"fma_per_clk_per_sm_ffma\": %.1f, \"fma_per_clk_per_sm_ffma2\": %.1f}\n",
           p.name, p.multiProcessorCount, clk / 1000, f1, ms1, f2, ms2,
           f1 * 1e12 / 2.0 / (p.multiProcessorCount * double(clk) * 1e3),
           f2 * 1e12 / 2.0 / (p.multiProcessorCount * double(clk) * 1e3));
    cudaFree(out);
    return cudaGetLastError() != cudaSuccess;
}

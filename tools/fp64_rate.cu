// fp64 throughput on this GPU: vector DFMA vs tensor-core DMMA (mma.sync m8n8k4 f64), 16 warps per SM, independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/fp64_rate tools/fp64_rate.cu && tools/build/fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) dfma_kernel(double *out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) dmma_kernel(double *out, int iters) {
    double c0[4], c1[4];
    for (int i = 0; i < 4; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) ffma_kernel(float *out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 512);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dfma_kernel<<<sms, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("DFMA  : %.1f FMA / clk / SM  (%.2f TFLOP/s)\n", 512.0 * 8 * iters / (ms * 1e-3 * khz * 1e3), 2.0 * sms * 512.0 * 8 * iters / (ms * 1e-3) / 1e12);
        cudaEventRecord(e0); dmma_kernel<<<sms, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("DMMA  : %.1f FMA / clk / SM  (%.2f TFLOP/s)\n", 16.0 * 4 * 256 * iters / (ms * 1e-3 * khz * 1e3), 2.0 * sms * 16.0 * 4 * 256 * iters / (ms * 1e-3) / 1e12);
        cudaEventRecord(e0); ffma_kernel<<<sms, 512>>>((float *)out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("FFMA  : %.1f FMA / clk / SM  (%.2f TFLOP/s)\n", 512.0 * 8 * iters / (ms * 1e-3 * khz * 1e3), 2.0 * sms * 512.0 * 8 * iters / (ms * 1e-3) / 1e12);
    }
    printf("SMs %d, clock %d MHz (attribute)\n", sms, khz / 1000);
    return 0;
}

// microlat.cu - dependent-chain latencies (cycles / op) of the instructions on the LAP step's critical path, one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/microlat tools/microlat.cu && tools/build/microlat
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(long long *out, double seed, unsigned useed) {
    const int lane = threadIdx.x & 31;
    long long t0, t1;
    // 1. DADD chain
    double a = seed + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + seed;
    t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    // 2. redux.sync min chain (each depends on the previous)
    unsigned u = useed + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = __reduce_min_sync(0xffffffffu, u + lane) + 1;
    t1 = clock64();
    if (lane == 0) out[1] = t1 - t0;
    // 3. shfl chain
    unsigned s = useed + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) s = __shfl_sync(0xffffffffu, s, (s + i) & 31) + 1;
    t1 = clock64();
    if (lane == 0) out[2] = t1 - t0;
    // 4. ballot chain
    unsigned b = useed + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) b = __ballot_sync(0xffffffffu, (b + lane + i) & 1) + lane;
    t1 = clock64();
    if (lane == 0) out[3] = t1 - t0;
    // 5. DSETP + select chain
    double c = seed + lane, d = seed * 2;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { c = (c < d) ? d + 1.0 : c - 1.0; }
    t1 = clock64();
    if (lane == 0) out[4] = t1 - t0;
    // 6. shared-memory load chain (pointer chasing)
    __shared__ int sm[64];
    sm[lane] = (lane * 7 + 1) & 31; sm[lane + 32] = lane;
    __syncwarp();
    int p = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = sm[p];
    t1 = clock64();
    if (lane == 0) out[5] = t1 - t0;
    // 7. integer add chain
    unsigned x = useed + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * 3 + i;
    t1 = clock64();
    if (lane == 0) out[6] = t1 - t0;
    // 8. DADD with an infinite operand
    double e = seed + lane, inf = seed > 0 ? __longlong_as_double(0x7ff0000000000000LL) : 1.0;
    double acc = 0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double r = (e + (double)i) - (-inf); acc = (r < acc) ? r : acc + 1.0; }
    t1 = clock64();
    if (lane == 0) out[7] = t1 - t0;
    if (a + u + s + b + c + p + x + acc == 12345.678) out[8] = 1;
}
int main() {
    long long *d, h[9];
    cudaMalloc(&d, sizeof(h));
    k<<<1, 32>>>(d, 1.5, 7u);
    k<<<1, 32>>>(d, 1.5, 7u);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[8] = {"DADD", "redux.sync.min (+1 add)", "shfl.idx (+add)", "ballot (+add)", "DSETP+select+DADD", "LDS pointer chase", "IMAD", "DADD(inf)+DSETP+sel"};
    for (int i = 0; i < 8; ++i) printf("%-28s %.1f cycles / iteration\n", names[i], (double)h[i] / N);
    return 0;
}

// Dependent-issue latencies on B200 (single warp, clock64): DFMA, DADD, SHFL(64-bit), LDS, MUFU.RCP64H+div, log.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x] = a + threadIdx.x;
    sm[threadIdx.x + 32] = b;
    __syncthreads();
    const int N = 512;
    double x = a + threadIdx.x * 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, a, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + b;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // SHFL 64-bit + DADD chain (one butterfly step)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // LDS dependent chain (pointer chase through index)
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double v = sm[idx]; idx = ((int) v) & 31; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    x += idx;
    // division chain
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = a / (x + b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // log chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = log(x + 2.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = sqrt(x + 2.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // IMAD/int add chain
    unsigned u = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { u += (u << 13) | (u >> 19); u ^= 0x9e3779b9u + i; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    out[threadIdx.x] = x + u;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, 0.999, 1e-3);
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char *names[8] = {"DFMA", "DADD", "SHFL64+DADD", "LDS+cvt", "DDIV+DADD", "log+DADD", "sqrt+DADD", "3-op int round"};
    for (int i = 0; i < 8; ++i) printf("%-16s %8.2f cycles per dependent op\n", names[i], h[i] / 512.0);
    return 0;
}

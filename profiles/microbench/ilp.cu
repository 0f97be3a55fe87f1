// FP64 issue behaviour on B200: cycles per DFMA for one CTA of W warps (warp w sits on SM sub-partition w % 4)
// running ILP independent dependent-chains per thread.  Tells whether a latency-bound warp can recover
// throughput through ILP (speculative proposals) or only through more warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, double a, double b, int slot) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    const int N = 2048;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N / 8; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[slot] = t1 - t0;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
    const int Ws[5] = {1, 4, 8, 12, 16};
    int slot = 0;
    for (int wi = 0; wi < 5; ++wi) {
        const int W = Ws[wi];
        for (int rep = 0; rep < 2; ++rep) {
            k<1><<<1, 32 * W>>>(out, cyc, 0.999, 1e-3, slot + 0);
            k<2><<<1, 32 * W>>>(out, cyc, 0.999, 1e-3, slot + 1);
            k<4><<<1, 32 * W>>>(out, cyc, 0.999, 1e-3, slot + 2);
            k<8><<<1, 32 * W>>>(out, cyc, 0.999, 1e-3, slot + 3);
        }
        slot += 4;
    }
    long long h[64];
    cudaMemcpy(h, cyc, 64 * 8, cudaMemcpyDeviceToHost);
    printf("cycles per DFMA instruction issued by one warp (2048 x ILP DFMAs per thread)\n");
    for (int wi = 0; wi < 5; ++wi) {
        printf("warps/CTA %2d (%.2f per sub-partition):", Ws[wi], Ws[wi] / 4.0);
        for (int j = 0; j < 4; ++j) printf("  ILP%d %6.2f", 1 << j, h[wi * 4 + j] / (2048.0 * (1 << j)));
        printf("\n");
    }
    return 0;
}

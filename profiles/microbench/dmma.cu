// FP64 tensor-core path on B200 (sm_100a): issue rate and latency of mma.sync.aligned.m8n8k4.f64 (SASS: DMMA),
// alone and next to DFMA streams.  Answers: is DMMA faster than DFMA per flop, does it run on a
// pipe of its own (so that a chain's quantile polynomials and its triangular matvec overlap), and how deep an
// accumulator chain can be before latency shows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ILP independent accumulator chains of DMMAs per warp, FMA extra DFMAs (8 independent chains) per DMMA
template <int ILP, int FMA>
__global__ void k(double *out, long long *cyc, double a, double b, int slot) {
    double c0[ILP], c1[ILP], x[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c0[i] = c1[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    const int N = 1024;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N / 4; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                dmma(c0[i], c1[i], a, b);
#pragma unroll
                for (int f = 0; f < FMA; ++f) x[f & 7] = fma(x[f & 7], a, b);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[slot] = t1 - t0;
}

template <int ILP, int FMA>
static double run(int warps, double *out, long long *cyc) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) k<ILP, FMA><<<1, 32 * warps>>>(out, cyc, 0.999, 1e-3, 0);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    return (double) h / (1024.0 * ILP);
}

// whole-GPU DMMA throughput
__global__ void __launch_bounds__(256) k_peak(double *out, long long iters, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c0[i] = c1[i] = threadIdx.x * 1e-9 + i;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 64 * 8);
    printf("cycles per DMMA (m8n8k4.f64 = 256 FMA = 8 warp-DFMAs) issued by one warp; one CTA of W warps\n");
    const int Ws[4] = {1, 4, 8, 16};
    for (int wi = 0; wi < 4; ++wi) {
        const int W = Ws[wi];
        printf("warps/CTA %2d (%.2f per sub-partition):  ILP1 %7.2f  ILP2 %7.2f  ILP4 %7.2f  ILP8 %7.2f\n", W, W / 4.0,
               run<1, 0>(W, out, cyc), run<2, 0>(W, out, cyc), run<4, 0>(W, out, cyc), run<8, 0>(W, out, cyc));
    }
    printf("\nDMMA (ILP4) interleaved with F independent DFMAs per DMMA, cycles per DMMA slot:\n");
    for (int wi = 0; wi < 3; ++wi) {
        const int W = Ws[wi];
        printf("warps/CTA %2d:  F0 %7.2f  F2 %7.2f  F4 %7.2f  F8 %7.2f  F16 %7.2f\n", W, run<4, 0>(W, out, cyc),
               run<4, 2>(W, out, cyc), run<4, 4>(W, out, cyc), run<4, 8>(W, out, cyc), run<4, 16>(W, out, cyc));
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const long long iters = 1 << 13;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_peak<<<sms * 8, 256>>>(out, iters, 0.999, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 256.0 * 8.0 * iters * (double) sms * 8 * 8 / (ms * 1e-3) / 1e12;
        if (rep && tf > best) best = tf;
    }
    printf("\nwhole-GPU DMMA throughput: %.2f TFLOP/s (%d SMs)\n", best, sms);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

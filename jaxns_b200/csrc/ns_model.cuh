// Cooperative evaluation of Model.forward for the registered likelihood families.
//
// Mapping (B200): one chain is owned by a group of G lanes (G = power of two <= 32, inside one
// warp); dimension j lives in lane (j % G), register slot (j / G), so a chain's D-vectors are
// spread over the group's registers (DPL = ceil(D / G) doubles per lane per vector) and every
// D-wide step (quantile transform, direction draw, cube bounds, matrix-vector product) runs
// lane-parallel with shuffle reductions.  G and DPL are compile-time so masks, strides and unrolls
// are constants.
//
// P proposals are evaluated per call ("batch"): the slice sampler's shrink sequence is known ahead
// of the likelihood values (see ns_slice.cuh), so P candidate points are pushed through the prior
// transform and the likelihood together.  A chain is one long dependency chain and config-2 sized
// problems only put ~2.7 warps on each SM sub-partition, so this instruction-level parallelism is
// what keeps the FP64 pipe busy.
//
// Likelihood parameters are staged once per CTA in shared memory.  The dense Gaussian factor is
// stored transposed (column j contiguous over rows) so that a warp reads consecutive rows
// conflict-free; for D <= 32 every lane keeps its row of L^-1 in registers instead.  The residuals
// r_j of the P proposals are exchanged through a per-chain shared-memory scratch [DP][P] read with
// broadcast loads.
//
// Reference: Model.forward (/root/reference/src/jaxns/framework/model.py:167-176) ->
// compute_log_likelihood (framework/ops.py:302-326, NaN -> -inf at :323-325) ->
// WrappedTFPDistribution._forward (framework/wrapped_tfp_distribution.py:77-84).
#pragma once
#include "../../include/nsb200.h"
#include "ns_math.cuh"

namespace nsb {

constexpr int kThreadsPerBlock = 128;

template <int G>
struct Grp {
    unsigned mask;  // lanes of this group inside the warp (compile-time constant for G == 32)
    int lane;       // lane index inside the group
    // convergent = every group of the warp executes the same control flow (team mode): shuffles and
    // syncs can then name the full warp, which keeps the mask a compile-time constant.
    __device__ __forceinline__ explicit Grp(bool convergent = false) {
        const int wl = threadIdx.x & 31;
        lane = wl & (G - 1);
        if (G == 32 || convergent) mask = 0xFFFFFFFFu;
        else mask = ((1u << (G & 31)) - 1u) << (wl & ~(G - 1));
    }
    __device__ __forceinline__ unsigned m() const { return G == 32 ? 0xFFFFFFFFu : mask; }
};

template <int G>
__device__ __forceinline__ void group_sync(const Grp<G> &g) {
    if (G > 1) __syncwarp(g.m());
}
template <int G>
__device__ __forceinline__ double group_sum(const Grp<G> &g, double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(g.m(), v, o);
    return v;
}
// P independent sums, butterfly steps interleaved (the shuffle latencies of the P chains overlap)
template <int G, int P>
__device__ __forceinline__ void group_sum_batch(const Grp<G> &g, double (&v)[P]) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
        double y[P];
#pragma unroll
        for (int p = 0; p < P; ++p) y[p] = __shfl_xor_sync(g.m(), v[p], o);
#pragma unroll
        for (int p = 0; p < P; ++p) v[p] += y[p];
    }
}
template <int G>
__device__ __forceinline__ double group_prod(const Grp<G> &g, double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v *= __shfl_xor_sync(g.m(), v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_min(const Grp<G> &g, double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(g.m(), v, o));
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(const Grp<G> &g, double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(g.m(), v, o));
    return v;
}

// Shared-memory image of the model, built once per CTA.
struct ModelSmem {
    int D, K, family, prior_kind;
    double *prior_a;  // [DP]
    double *prior_b;  // [DP]
    double *params;   // family specific, see stage_model()
    const double *dense_global;  // row-major Linv in global memory when the factor does not fit in smem
};

// Largest dense factor (in doubles) staged in shared memory; beyond it the kernel reads the
// row-major factor from global memory (L1/L2 resident) instead.
constexpr size_t kDenseSmemMaxDoubles = 20 * 1024;  // 160 KB

__host__ __device__ inline bool dense_in_smem(int D, int DP) { return (size_t) D * DP <= kDenseSmemMaxDoubles; }
// D <= 32 with one lane per dimension: the factor row lives in registers, nothing is staged.
__host__ __device__ constexpr bool dense_in_regs(int G, int DPL) { return G == 32 && DPL == 1; }

// Doubles of shared memory the staged model needs.
// row_smem: stage the factor in shared memory even where dense_in_regs() would keep it in registers (warp teams:
// two warps per chain only stay resident in one wave at <= 88 registers per thread, the 64-register row has to go).
__host__ __device__ inline size_t model_smem_doubles(int family, int D, int G, int DPL, int K, bool row_smem = false) {
    const int DP = G * DPL;
    size_t n = 2 * (size_t) DP;
    switch (family) {
        case NSB200_FAM_GAUSS_DENSE:
            n += 1 + DP + (((!dense_in_regs(G, DPL) || row_smem) && dense_in_smem(D, DP)) ? (size_t) D * DP : 0);  // c, mu, LT[D][DP]
            break;
        case NSB200_FAM_GAUSS_MIX_DIAG: n += (size_t) K * (1 + 2 * (size_t) DP); break;  // logc, mean[DP], inv[DP]
        case NSB200_FAM_SHELLS: n += (size_t) K * (3 + (size_t) DP); break;              // w, r, lognorm, c[DP]
        default: break;
    }
    return n;
}

// Cooperative (whole CTA) staging of the model into shared memory.  Padded dimensions get neutral
// values.  Must be followed by __syncthreads().
template <int G, int DPL, bool RS = false>
__device__ inline void stage_model(const NsModelDesc &m, double *smem, ModelSmem &out) {
    constexpr int DP = G * DPL;
    const int D = m.D;
    out.D = D;
    out.K = m.K;
    out.family = m.family;
    out.prior_kind = m.prior_kind;
    out.prior_a = smem;
    out.prior_b = smem + DP;
    out.params = smem + 2 * DP;
    out.dense_global = nullptr;
    for (int j = threadIdx.x; j < DP; j += blockDim.x) {
        out.prior_a[j] = (j < D) ? m.prior_a[j] : 0.0;
        out.prior_b[j] = (j < D) ? m.prior_b[j] : 0.0;
    }
    double *P = out.params;
    const double *src = m.params;
    switch (m.family) {
        case NSB200_FAM_GAUSS_DENSE: {
            // src = [c, mu[D], Linv[D*D] row-major]; dst = [c, mu[DP], LT[j][i] = Linv[i][j]]
            if (threadIdx.x == 0) P[0] = src[0];
            for (int j = threadIdx.x; j < DP; j += blockDim.x) P[1 + j] = (j < D) ? src[1 + j] : 0.0;
            if (dense_in_regs(G, DPL) && !RS) {
                out.dense_global = src + 1 + D;  // rows are pulled into registers by the caller
            } else if (dense_in_smem(D, DP)) {
                double *LT = P + 1 + DP;
                for (int e = threadIdx.x; e < D * DP; e += blockDim.x) {
                    int j = e / DP, i = e - j * DP;
                    LT[e] = (i < D && j <= i) ? src[1 + D + (size_t) i * D + j] : 0.0;
                }
            } else {
                out.dense_global = src + 1 + D;
            }
            break;
        }
        case NSB200_FAM_GAUSS_MIX_DIAG: {
            for (int e = threadIdx.x; e < m.K * (1 + 2 * DP); e += blockDim.x) {
                int k = e / (1 + 2 * DP), o = e - k * (1 + 2 * DP);
                const double *sk = src + (size_t) k * (1 + 2 * D);
                double v;
                if (o == 0) v = sk[0];
                else if (o <= DP) { int j = o - 1; v = (j < D) ? sk[1 + j] : 0.0; }
                else { int j = o - 1 - DP; v = (j < D) ? sk[1 + D + j] : 0.0; }
                P[e] = v;
            }
            break;
        }
        case NSB200_FAM_SHELLS: {
            for (int e = threadIdx.x; e < m.K * (3 + DP); e += blockDim.x) {
                int k = e / (3 + DP), o = e - k * (3 + DP);
                const double *sk = src + (size_t) k * (2 + D);
                double v;
                if (o < 2) v = sk[o];
                else if (o == 2) v = log(sqrt(2.0 * 3.14159265358979323846 * (sk[0] * sk[0])));
                else { int j = o - 3; v = (j < D) ? sk[2 + j] : 0.0; }
                P[e] = v;
            }
            break;
        }
        default: break;
    }
}

// Matvec of the D <= 32 dense factor on the FP64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) INSIDE the
// lane-per-dimension kernel: the P proposals of a round are columns 0..P-1 of the B operand (the other columns are
// zero), L^-1 is 20 lower-triangular 8x4 A tiles per lane (40 registers instead of the 64-register row), the
// residuals reach the B layout with one shuffle per k tile and proposal instead of a shared-memory exchange, and
// ||z||^2 needs a 3-step butterfly instead of 5.  This is a LATENCY trade: 20 DMMAs occupy the FP64 pipe for 320
// cycles where 32 P DFMAs need 64 P, but the dependent chain of a round (exchange -> 32-deep FMA row -> butterfly)
// gets shorter, and the round's latency is what bounds this kernel.  MEASURED (config 2, -DNSB_LANE_MMA=1):
// 140 registers instead of 166, same results, slice kernel 0.617 ms instead of 0.537 ms -- with 2.7 warps per
// sub-partition the 320 pipe cycles per round and warp contend (FP64 pipe 40 % -> ~66 %) and cost more than the
// shorter chain saves.  Off by default; kept because it is the cheapest way to re-measure on other shapes.
#ifndef NSB_LANE_MMA
#define NSB_LANE_MMA 0
#endif
constexpr bool kLaneMma = NSB_LANE_MMA != 0;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Sum over the 32 lanes of a warp through ONE FP64 tensor-core instruction per value instead of a 5-step butterfly:
// with A = ones(8x4) and B[k][n] = the value of lane 4 n + k, C[m][n] = sum over the 4 lanes of group n for every row
// m; lane (g, t) holds C[g][2 t] and C[g][2 t + 1], so c0 + c1 is the sum over 8 lanes and two shuffle steps over t
// finish it (DMMA 26 cycles + add + 2 x (shuffle + add) against 5 x (shuffle + add)).  Measured on the same box
// (profiles/r2/regcap_ab_r2.txt): one launch over 200 / 1600 / 25600 chains of config 2 0.510 -> 0.489 / 0.697 -> 0.691 /
// 8.04 -> 7.92 ms, a whole run 95.2 -> 94.5 ms.  -DNSB_DMMA_REDUCE=0 restores the butterfly.
#ifndef NSB_DMMA_REDUCE
#define NSB_DMMA_REDUCE 1
#endif
constexpr bool kDmmaReduce = NSB_DMMA_REDUCE != 0;

template <int P>
__device__ __forceinline__ void warp_sum_dmma(double (&v)[P]) {
    double s[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        double c0 = 0.0, c1 = 0.0;
        dmma_m8n8k4(c0, c1, 1.0, v[p]);
        s[p] = c0 + c1;
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        double y[P];
#pragma unroll
        for (int p = 0; p < P; ++p) y[p] = __shfl_xor_sync(0xFFFFFFFFu, s[p], o);
#pragma unroll
        for (int p = 0; p < P; ++p) s[p] += y[p];
    }
#pragma unroll
    for (int p = 0; p < P; ++p) v[p] = s[p];
}

// Per-thread registers of the dense factor (only meaningful for G == 32, DPL == 1): the lane's row of L^-1, or
// (kLaneMma) its elements of the 20 A tiles on or below the diagonal: tile (b, s), s <= 2 b + 1, holds
// Linv[8 b + lane / 4][4 s + lane % 4].
template <int G, int DPL, bool RS = false>
struct DenseRow {
    static constexpr bool kRegs = dense_in_regs(G, DPL) && !RS;
    static constexpr int N = kRegs ? (kLaneMma ? 20 : 32) : 1;
    double v[N];
    __device__ __forceinline__ void load(const ModelSmem &sm, int lane) {
        if (kRegs && kLaneMma) {
            int ai = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
#pragma unroll
                for (int s = 0; s < 2 * b + 2; ++s) {
                    const int row = 8 * b + (lane >> 2), col = 4 * s + (lane & 3);
                    v[ai++] = (sm.family == NSB200_FAM_GAUSS_DENSE && row < sm.D && col < sm.D && col <= row)
                                  ? __ldg(sm.dense_global + (size_t) row * sm.D + col)
                                  : 0.0;
                }
            }
        } else if (kRegs) {
#pragma unroll
            for (int j = 0; j < N; ++j)
                v[j] = (sm.family == NSB200_FAM_GAUSS_DENSE && lane < sm.D && j <= lane)
                           ? __ldg(sm.dense_global + (size_t) lane * sm.D + j)
                           : 0.0;
        }
    }
};

// Prior quantile transform of P points for this lane's dimensions.
template <int G, int DPL, int P>
__device__ __forceinline__ void transform_dims(const ModelSmem &sm, const Grp<G> &g, const double (&u)[P][DPL],
                                               double (&X)[P][DPL]) {
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        const double a = sm.prior_a[j], b = sm.prior_b[j];
        if (sm.prior_kind == NSB200_PRIOR_UNIFORM) {
#pragma unroll
            for (int p = 0; p < P; ++p) X[p][s] = u[p][s] * b + a;
        } else if (P > 1 && G >= 16) {
            double up[P], zp[P];
#pragma unroll
            for (int p = 0; p < P; ++p) up[p] = u[p][s];
            ndtri_batch<P>(up, g.m(), zp);
#pragma unroll
            for (int p = 0; p < P; ++p) X[p][s] = zp[p] * b + a;
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) X[p][s] = ndtri(u[p][s], g.m()) * b + a;  // padded dims: u = 0.5, b = 0
        }
    }
}

// log-likelihood of the P transformed points held across the group.  `scratch` = DP*P doubles of
// shared memory private to the chain.  Every lane of the group gets the same values.
template <int G, int DPL, int P, int FAM, bool RS = false>
__device__ __forceinline__ void loglik_group(const ModelSmem &sm, const Grp<G> &g, const DenseRow<G, DPL, RS> &row,
                                             const double (&X)[P][DPL], double *scratch, double (&out)[P]) {
    constexpr int DP = G * DPL;
    const int D = sm.D;
    const double *Pm = sm.params;
    const double kNan = __longlong_as_double(0x7FF8000000000000ll);
    switch (FAM) {  // compile-time: the kernels dispatch on the family once, outside the chain loop
        case NSB200_FAM_GAUSS_DENSE: {
            const double *mu = Pm + 1;
            if (DenseRow<G, DPL, RS>::kRegs && kLaneMma && P <= 8) {
                // residual of dimension `lane` for each proposal, then B tile s: lane (g, t) needs r[4 s + t] of
                // proposal g (columns >= P are zero); the 8 tiles are shared by the 4 row blocks
                const int lane = g.lane, gq = lane >> 2, tq = lane & 3;
                double r[P];
                const double m = mu[lane];
#pragma unroll
                for (int p = 0; p < P; ++p) r[p] = (lane < D) ? X[p][0] - m : 0.0;
                double bt[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    double bv = 0.0;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const double y = __shfl_sync(0xFFFFFFFFu, r[p], 4 * s + tq);
                        bv = (gq == p) ? y : bv;
                    }
                    bt[s] = bv;
                }
                double sq0 = 0.0, sq1 = 0.0;
                int ai = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int s = 0; s < 2 * b + 2; ++s) dmma_m8n8k4(c0, c1, row.v[ai++], bt[s]);
                    sq0 = fma(c0, c0, sq0);
                    sq1 = fma(c1, c1, sq1);
                }
                // C fragment: rows g (+ 8 b), columns 2 t and 2 t + 1 -> sum over the rows = lanes with equal t
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    const double y0 = __shfl_xor_sync(0xFFFFFFFFu, sq0, o), y1 = __shfl_xor_sync(0xFFFFFFFFu, sq1, o);
                    sq0 += y0;
                    sq1 += y1;
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const double qv = __shfl_sync(0xFFFFFFFFu, (p & 1) ? sq1 : sq0, p >> 1);  // column p: lanes with t = p / 2
                    out[p] = Pm[0] - 0.5 * qv;
                }
                break;
            }
            // residuals r_j^(p) -> scratch[j][p]
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int j = s * G + g.lane;
                const double m = mu[j];
#pragma unroll
                for (int p = 0; p < P; ++p) scratch[j * P + p] = (j < D) ? X[p][s] - m : 0.0;
            }
            group_sync(g);
            double q[P];
#pragma unroll
            for (int p = 0; p < P; ++p) q[p] = 0.0;
            if (dense_in_regs(G, DPL) && !RS) {
                // z_lane = sum_j Linv[lane][j] r_j with the row in registers; two partial sums per
                // proposal shorten the dependent FMA chain.
                constexpr int NR = DenseRow<G, DPL, RS>::N;
#ifndef NSB_ACC
#define NSB_ACC 4
#endif
                constexpr int NA = NR >= NSB_ACC ? NSB_ACC : 1;  // partial sums: 8-deep FMA chains instead of 32
                double z[NA][P];
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int p = 0; p < P; ++p) z[a][p] = 0.0;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
#pragma unroll
                    for (int p = 0; p < P; ++p) z[j % NA][p] = fma(row.v[j], scratch[j * P + p], z[j % NA][p]);
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    double zz = z[0][p];
                    if (NA == 4) zz = (z[0][p] + z[1][p]) + (z[2][p] + z[3][p]);
                    if (NA == 8) zz = ((z[0][p] + z[1][p]) + (z[2][p] + z[3][p])) + ((z[4][p] + z[5][p]) + (z[6][p] + z[7][p]));
                    q[p] = zz * zz;
                }
            } else if (sm.dense_global) {
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int i = s * G + g.lane;
                    if (i < D) {
                        const double *rw = sm.dense_global + (size_t) i * D;
                        double acc[P];
#pragma unroll
                        for (int p = 0; p < P; ++p) acc[p] = 0.0;
                        for (int jj = 0; jj <= i; ++jj) {
                            const double l = __ldg(rw + jj);
#pragma unroll
                            for (int p = 0; p < P; ++p) acc[p] = fma(l, scratch[jj * P + p], acc[p]);
                        }
#pragma unroll
                        for (int p = 0; p < P; ++p) q[p] = fma(acc[p], acc[p], q[p]);
                    }
                }
            } else {
                const double *LT = Pm + 1 + DP;
                // rows of slot s end at (s+1)G-1: column loop stops there (lower-triangular factor)
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jend = min(D, (s + 1) * G);
                    const double *col = LT + s * G + g.lane;
                    double a0[P], a1[P];
#pragma unroll
                    for (int p = 0; p < P; ++p) a0[p] = a1[p] = 0.0;
                    int jj = 0;
                    for (; jj + 1 < jend; jj += 2) {
                        const double l0 = col[jj * DP], l1 = col[(jj + 1) * DP];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            a0[p] = fma(l0, scratch[jj * P + p], a0[p]);
                            a1[p] = fma(l1, scratch[(jj + 1) * P + p], a1[p]);
                        }
                    }
                    if (jj < jend) {
                        const double l0 = col[jj * DP];
#pragma unroll
                        for (int p = 0; p < P; ++p) a0[p] = fma(l0, scratch[jj * P + p], a0[p]);
                    }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const double z = a0[p] + a1[p];
                        q[p] = fma(z, z, q[p]);
                    }
                }
            }
            if (kDmmaReduce && G == 32) {
                warp_sum_dmma<P>(q);
#pragma unroll
                for (int p = 0; p < P; ++p) out[p] = Pm[0] - 0.5 * q[p];
            } else if (P > 1) {
                group_sum_batch<G, P>(g, q);
#pragma unroll
                for (int p = 0; p < P; ++p) out[p] = Pm[0] - 0.5 * q[p];
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p) out[p] = Pm[0] - 0.5 * group_sum(g, q[p]);
            }
            break;
        }
        case NSB200_FAM_GAUSS_MIX_DIAG: {
            for (int k = 0; k < sm.K; ++k) {
                const double *pk = Pm + (size_t) k * (1 + 2 * DP);
                double q[P];
#pragma unroll
                for (int p = 0; p < P; ++p) q[p] = 0.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int j = s * G + g.lane;
                    const double mean = pk[1 + j], inv = pk[1 + DP + j];
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const double zz = (X[p][s] - mean) * inv;
                        q[p] = fma(zz, zz, q[p]);
                    }
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const double gk = pk[0] - 0.5 * group_sum(g, q[p]);
                    out[p] = (k == 0) ? gk : logaddexp(out[p], gk);
                }
            }
            break;
        }
        case NSB200_FAM_EGGBOX: {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                double y = 1.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int j = s * G + g.lane;
                    if (j < D) y *= cos(0.5 * X[p][s]);
                }
                y = 2.0 + group_prod(g, y);
                const double y2 = y * y;
                out[p] = y2 * y2 * y;
            }
            break;
        }
        case NSB200_FAM_ROSENBROCK: {
#pragma unroll
            for (int s = 0; s < DPL; ++s)
#pragma unroll
                for (int p = 0; p < P; ++p) scratch[(s * G + g.lane) * P + p] = X[p][s];
            group_sync(g);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                double y = 0.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int j = s * G + g.lane;
                    if (j < D - 1) {
                        const double a = scratch[(j + 1) * P + p] - X[p][s] * X[p][s];
                        const double b = 1.0 - X[p][s];
                        y += 100.0 * (a * a) + b * b;
                    }
                }
                out[p] = -group_sum(g, y);
            }
            break;
        }
        case NSB200_FAM_SHELLS: {
            for (int k = 0; k < sm.K; ++k) {
                const double *pk = Pm + (size_t) k * (3 + DP);
                const double w = pk[0], rad = pk[1], lognorm = pk[2];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    double ssq = 0.0;
#pragma unroll
                    for (int s = 0; s < DPL; ++s) {
                        const int j = s * G + g.lane;
                        const double dl = (j < D) ? X[p][s] - pk[3 + j] : 0.0;
                        ssq = fma(dl, dl, ssq);
                    }
                    ssq = group_sum(g, ssq);
                    const double e = sqrt(ssq) - rad;
                    const double gk = -0.5 * (e * e) / (w * w) - lognorm;
                    out[p] = (k == 0) ? gk : logaddexp(out[p], gk);
                }
            }
            break;
        }
        default:
#pragma unroll
            for (int p = 0; p < P; ++p) out[p] = kNan;
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (out[p] != out[p]) out[p] = -__longlong_as_double(0x7FF0000000000000ll);  // ops.py:323-325
    // the scratch is rewritten by the next call: make sure every lane is done reading it
    group_sync(g);
}

// Model.forward at P U-space points held across the group.
template <int G, int DPL, int P, int FAM, bool RS = false>
__device__ __forceinline__ void forward_group(const ModelSmem &sm, const Grp<G> &g, const DenseRow<G, DPL, RS> &row,
                                              const double (&u)[P][DPL], double *scratch, double (&out)[P]) {
    double X[P][DPL];
    transform_dims<G, DPL, P>(sm, g, u, X);
    loglik_group<G, DPL, P, FAM, RS>(sm, g, row, X, scratch, out);
}

// Runtime family -> compile-time template argument, once per kernel.
#ifdef NSB_FAST_BUILD
#define NSB_FAMILY_SWITCH(FAMILY, ...)                                                          \
    switch (FAMILY) {                                                                           \
        case NSB200_FAM_GAUSS_DENSE: { constexpr int kFam = NSB200_FAM_GAUSS_DENSE; __VA_ARGS__; } break;       \
        default: break;                                                                         \
    }
#else
#define NSB_FAMILY_SWITCH(FAMILY, ...)                                                          \
    switch (FAMILY) {                                                                           \
        case NSB200_FAM_GAUSS_DENSE: { constexpr int kFam = NSB200_FAM_GAUSS_DENSE; __VA_ARGS__; } break;       \
        case NSB200_FAM_GAUSS_MIX_DIAG: { constexpr int kFam = NSB200_FAM_GAUSS_MIX_DIAG; __VA_ARGS__; } break; \
        case NSB200_FAM_EGGBOX: { constexpr int kFam = NSB200_FAM_EGGBOX; __VA_ARGS__; } break;                 \
        case NSB200_FAM_ROSENBROCK: { constexpr int kFam = NSB200_FAM_ROSENBROCK; __VA_ARGS__; } break;         \
        case NSB200_FAM_SHELLS: { constexpr int kFam = NSB200_FAM_SHELLS; __VA_ARGS__; } break;                 \
        default: break;                                                                         \
    }
#endif

}  // namespace nsb

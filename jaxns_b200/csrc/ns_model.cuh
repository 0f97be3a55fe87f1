// Cooperative evaluation of Model.forward for the registered likelihood families.
//
// Mapping (B200): one chain is owned by a group of G lanes (G = power of two <= 32, inside one
// warp); dimension j lives in lane (j % G), register slot (j / G), so a chain's D-vectors are
// spread over the group's registers (DPL = ceil(D / G) doubles per lane per vector) and every
// D-wide step (quantile transform, direction draw, cube bounds, matrix-vector product) runs
// lane-parallel with shuffle reductions.  Likelihood parameters are staged once per CTA in shared
// memory; the dense Gaussian factor is stored transposed (column j contiguous over rows) so that a
// warp reads consecutive rows conflict-free while r_j is a broadcast load.
//
// Reference: Model.forward (/root/reference/src/jaxns/framework/model.py:167-176) ->
// compute_log_likelihood (framework/ops.py:302-326, NaN -> -inf at :323-325) ->
// WrappedTFPDistribution._forward (framework/wrapped_tfp_distribution.py:77-84).
#pragma once
#include "../../include/nsb200.h"
#include "ns_math.cuh"

namespace nsb {

constexpr int kThreadsPerBlock = 128;

struct Grp {
    unsigned mask;  // lanes of this group inside the warp
    int lane;       // lane index inside the group
    int G;          // group size
};

__device__ __forceinline__ Grp make_group(int G) {
    Grp g;
    g.G = G;
    const int wl = threadIdx.x & 31;
    g.lane = wl & (G - 1);
    const unsigned base = (G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u);
    g.mask = base << (wl & ~(G - 1));
    return g;
}

__device__ __forceinline__ void group_sync(const Grp &g) { __syncwarp(g.mask); }

__device__ __forceinline__ double group_sum(const Grp &g, double v) {
    for (int o = g.G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(g.mask, v, o);
    return v;
}
__device__ __forceinline__ double group_prod(const Grp &g, double v) {
    for (int o = g.G >> 1; o > 0; o >>= 1) v *= __shfl_xor_sync(g.mask, v, o);
    return v;
}
__device__ __forceinline__ double group_min(const Grp &g, double v) {
    for (int o = g.G >> 1; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(g.mask, v, o));
    return v;
}
__device__ __forceinline__ double group_max(const Grp &g, double v) {
    for (int o = g.G >> 1; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(g.mask, v, o));
    return v;
}
__device__ __forceinline__ long long group_sum_ll(const Grp &g, long long v) {
    for (int o = g.G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(g.mask, v, o);
    return v;
}

// Shared-memory image of the model, built once per CTA.
struct ModelSmem {
    int D, DP, K, family, prior_kind;
    double *prior_a;  // [DP]
    double *prior_b;  // [DP]
    double *params;   // family specific, see stage_model()
    const double *dense_global;  // row-major Linv in global memory when the factor does not fit in smem
};

// Largest dense factor (in doubles) staged in shared memory; beyond it the kernel reads the
// row-major factor from global memory (L1/L2 resident) instead.
constexpr size_t kDenseSmemMaxDoubles = 20 * 1024;  // 160 KB

__host__ __device__ inline bool dense_in_smem(int D, int DP) { return (size_t) D * DP <= kDenseSmemMaxDoubles; }

// Doubles of shared memory the staged model needs.
__host__ __device__ inline size_t model_smem_doubles(int family, int D, int DP, int K) {
    size_t n = 2 * (size_t) DP;
    switch (family) {
        case NSB200_FAM_GAUSS_DENSE: n += 1 + DP + (dense_in_smem(D, DP) ? (size_t) D * DP : 0); break;  // c, mu[DP], LT[D][DP]
        case NSB200_FAM_GAUSS_MIX_DIAG: n += (size_t) K * (1 + 2 * (size_t) DP); break;  // logc, mean[DP], inv[DP]
        case NSB200_FAM_SHELLS: n += (size_t) K * (2 + (size_t) DP); break;           // w, r, c[DP]
        default: break;
    }
    return n;
}

// Cooperative (whole CTA) staging of the model into shared memory.  Padded dimensions get neutral
// values.  Must be followed by __syncthreads().
__device__ inline void stage_model(const NsModelDesc &m, int DP, double *smem, ModelSmem &out) {
    const int D = m.D;
    out.D = D;
    out.DP = DP;
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
            if (dense_in_smem(D, DP)) {
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
            for (int e = threadIdx.x; e < m.K * (2 + DP); e += blockDim.x) {
                int k = e / (2 + DP), o = e - k * (2 + DP);
                const double *sk = src + (size_t) k * (2 + D);
                double v;
                if (o < 2) v = sk[o];
                else { int j = o - 2; v = (j < D) ? sk[2 + j] : 0.0; }
                P[e] = v;
            }
            break;
        }
        default: break;
    }
}

// Prior quantile transform for this lane's dimensions.
template <int DPL>
__device__ __forceinline__ void transform_dims(const ModelSmem &sm, const Grp &g, const double (&u)[DPL],
                                               double (&X)[DPL]) {
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * g.G + g.lane;
        const double a = sm.prior_a[j], b = sm.prior_b[j];
        if (sm.prior_kind == NSB200_PRIOR_UNIFORM) X[s] = u[s] * b + a;
        else X[s] = (j < sm.D) ? ndtri(u[s]) * b + a : 0.0;
    }
}

// log-likelihood of the transformed point held across the group.  `scratch` = DP doubles of shared
// memory private to the chain.  Returns the same value in every lane of the group.
template <int DPL>
__device__ __forceinline__ double loglik_group(const ModelSmem &sm, const Grp &g, const double (&X)[DPL],
                                               double *scratch) {
    const int D = sm.D, DP = sm.DP, G = g.G;
    const double *P = sm.params;
    double r;
    switch (sm.family) {
        case NSB200_FAM_GAUSS_DENSE: {
            const double *mu = P + 1;
            const double *LT = P + 1 + DP;
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int j = s * G + g.lane;
                scratch[j] = (j < D) ? X[s] - mu[j] : 0.0;
            }
            group_sync(g);
            double z[DPL];
#pragma unroll
            for (int s = 0; s < DPL; ++s) z[s] = 0.0;
            // z_i = sum_{j<=i} Linv[i][j] r_j ; rows of slot s end at (s+1)G-1
            if (sm.dense_global) {
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int i = s * G + g.lane;
                    if (i < D) {
                        const double *row = sm.dense_global + (size_t) i * D;
                        double acc = 0.0;
                        for (int jj = 0; jj <= i; ++jj) acc = fma(__ldg(row + jj), scratch[jj], acc);
                        z[s] = acc;
                    }
                }
            } else
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int jend = min(D, (s + 1) * G);
                const double *col = LT + s * G + g.lane;
                double acc0 = 0.0, acc1 = 0.0;
                int jj = 0;
                for (; jj + 1 < jend; jj += 2) {
                    acc0 = fma(col[(size_t) jj * DP], scratch[jj], acc0);
                    acc1 = fma(col[(size_t) (jj + 1) * DP], scratch[jj + 1], acc1);
                }
                if (jj < jend) acc0 = fma(col[(size_t) jj * DP], scratch[jj], acc0);
                z[s] = acc0 + acc1;
            }
            double q = 0.0;
#pragma unroll
            for (int s = 0; s < DPL; ++s) q = fma(z[s], z[s], q);
            q = group_sum(g, q);
            r = P[0] - 0.5 * q;
            break;
        }
        case NSB200_FAM_GAUSS_MIX_DIAG: {
            r = 0.0;
            for (int k = 0; k < sm.K; ++k) {
                const double *pk = P + (size_t) k * (1 + 2 * DP);
                double q = 0.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int j = s * G + g.lane;
                    double zz = (X[s] - pk[1 + j]) * pk[1 + DP + j];
                    q = fma(zz, zz, q);
                }
                q = group_sum(g, q);
                double gk = pk[0] - 0.5 * q;
                r = (k == 0) ? gk : logaddexp(r, gk);
            }
            break;
        }
        case NSB200_FAM_EGGBOX: {
            double y = 1.0;
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int j = s * G + g.lane;
                if (j < D) y *= cos(0.5 * X[s]);
            }
            y = 2.0 + group_prod(g, y);
            double y2 = y * y;
            r = y2 * y2 * y;
            break;
        }
        case NSB200_FAM_ROSENBROCK: {
#pragma unroll
            for (int s = 0; s < DPL; ++s) scratch[s * G + g.lane] = X[s];
            group_sync(g);
            double y = 0.0;
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int j = s * G + g.lane;
                if (j < D - 1) {
                    double a = scratch[j + 1] - X[s] * X[s];
                    double b = 1.0 - X[s];
                    y += 100.0 * (a * a) + b * b;
                }
            }
            r = -group_sum(g, y);
            break;
        }
        case NSB200_FAM_SHELLS: {
            r = 0.0;
            for (int k = 0; k < sm.K; ++k) {
                const double *pk = P + (size_t) k * (2 + DP);
                double ssq = 0.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int j = s * G + g.lane;
                    double dl = (j < D) ? X[s] - pk[2 + j] : 0.0;
                    ssq = fma(dl, dl, ssq);
                }
                ssq = group_sum(g, ssq);
                const double w = pk[0], rad = pk[1];
                double e = sqrt(ssq) - rad;
                double gk = -0.5 * (e * e) / (w * w) - log(sqrt(2.0 * 3.14159265358979323846 * (w * w)));
                r = (k == 0) ? gk : logaddexp(r, gk);
            }
            break;
        }
        default:
            r = __longlong_as_double(0x7FF8000000000000ll);
    }
    if (r != r) r = -__longlong_as_double(0x7FF0000000000000ll);  // ops.py:323-325
    return r;
}

// Model.forward at the U-space point held across the group.
template <int DPL>
__device__ __forceinline__ double forward_group(const ModelSmem &sm, const Grp &g, const double (&u)[DPL],
                                                double *scratch) {
    double X[DPL];
    transform_dims<DPL>(sm, g, u, X);
    return loglik_group<DPL>(sm, g, X, scratch);
}

}  // namespace nsb

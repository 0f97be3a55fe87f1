// Batched constrained samplers: the fused slice-chain kernel (Threefry + seed choice + direction +
// cube bracket + shrink loop + prior transform + likelihood, all in one launch), the prior-draw
// kernel for the initial live set, the uniform rejection sampler and vmap(Model.forward).
//
// Reference: get_samples (/root/reference/src/jaxns/nested_samplers/sharded/sharded_static.py:88-129),
// BaseAbstractMarkovSampler._get_sample (samplers/bases.py:63-75),
// UniDimSliceSampler.get_seed_point / get_sample_from_seed / _new_proposal
// (samplers/uni_slice_sampler.py:343-441, :114-273), resample_indicies (internals/random.py:55-60),
// _single_uniform_sample (nested_samplers/common/uniform_sample.py:12-60),
// UniformSampler._get_sample (samplers/uniform_samplers.py:42-85).
#pragma once
#include "ns_model.cuh"

namespace nsb {

// Number of uniforms per slice drawn ahead of the data (the key stream of a slice does not depend
// on the likelihood values, so lane L of a group precomputes the stream of slice base+L while its
// neighbours do the same for theirs; the shrink loop then only reads them).
constexpr int kPre = 8;

struct SliceArgs {
    NsModelDesc model;
    Key key;                // get_samples key: chain keys = split(key, m)
    const double *contour;  // device scalar
    const double *live_U;   // [N, D]
    const double *live_logL;
    const double *seed_table;
    double *out_U;
    double *out_logL;
    long long *out_nevals;
    double *ph_U;
    double *ph_logL;
    long long N;
    long long chain_begin, chain_end;
    int S, k, midpoint, G;
    // optional packed-row output for the multi-GPU all-gather (row = [U[D], logL, nevals, k x (U[D], logL)])
    double *packed;
    long long packed_row_doubles;
};

// jnp.linspace(0.5, 1., S)[j]
__device__ __forceinline__ double alpha_schedule(int j, int S) {
    if (S == 1) return 0.5;
    const int div = S - 1;
    if (j == div) return 1.0;
    const double step = (double) j / (double) div;
    return 0.5 * (1.0 - step) + 1.0 * step;
}

// _sample_direction: d = normal(key, (D,)); d /= ||d||
template <int DPL>
__device__ __forceinline__ void sample_direction(const Grp &g, int D, Key key, double (&d)[DPL]) {
    if (D == 1) {
#pragma unroll
        for (int s = 0; s < DPL; ++s) d[s] = (s == 0 && g.lane == 0) ? 1.0 : 0.0;
        return;
    }
    double ss = 0.0;
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * g.G + g.lane;
        d[s] = (j < D) ? normal_from_bits(bits64(key, (uint64_t) j)) : 0.0;
        ss = fma(d[s], d[s], ss);
    }
    const double nrm = sqrt(group_sum(g, ss));
#pragma unroll
    for (int s = 0; s < DPL; ++s) d[s] = d[s] / nrm;
}

// _slice_bounds: intersection of the line U0 + t d with the unit cube.
template <int DPL>
__device__ __forceinline__ void slice_bounds(const Grp &g, int D, const double (&U0)[DPL], const double (&d)[DPL],
                                             double &left, double &right) {
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    double r = kInf, l = -kInf;
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * g.G + g.lane;
        if (j < D) {
            const double t1 = (1.0 - U0[s]) / d[s];
            const double t0 = -U0[s] / d[s];
            if (t1 >= 0.0) r = fmin(r, t1);
            if (t1 <= 0.0) l = fmax(l, t1);
            if (t0 >= 0.0) r = fmin(r, t0);
            if (t0 <= 0.0) l = fmax(l, t0);
        }
    }
    right = group_min(g, r);
    left = group_max(g, l);
}

// Seed choice: first index with log_L > contour, then lower_bound in the logaddexp table.
__device__ __forceinline__ long long seed_index(const double *live_logL, const double *ctab, long long N,
                                                double contour, double u) {
    long long lo = 0, hi = N;
    while (lo < hi) {
        long long mid = lo + ((hi - lo) >> 1);
        if (live_logL[mid] > contour) hi = mid; else lo = mid + 1;
    }
    const long long j0 = lo, nsat = N - j0;
    if (nsat == 0) return 0;
    const double log_r = ctab[nsat - 1] + log(1.0 - u);
    lo = 0;
    hi = nsat;
    while (lo < hi) {
        long long mid = lo + ((hi - lo) >> 1);
        if (ctab[mid] < log_r) lo = mid + 1; else hi = mid;
    }
    return j0 + lo;
}

// Dynamic shared memory layout of the sampler kernels:
//   [model image][per chain: scratch[DP] | pre_u[G][kPre] | pre_keys[G][4 x u32]]
__host__ __device__ inline size_t chain_smem_doubles(int DP, int G) { return (size_t) DP + (size_t) G * (kPre + 2); }

template <int DPL>
__device__ __forceinline__ void slice_chains_body(const SliceArgs &a, double *smem) {
    const int G = a.G, DP = G * DPL, D = a.model.D;
    ModelSmem sm;
    stage_model(a.model, DP, smem, sm);
    __syncthreads();
    const Grp g = make_group(G);
    const int chains_per_block = kThreadsPerBlock / G;
    const int local_chain = threadIdx.x / G;
    const long long chain = a.chain_begin + (long long) blockIdx.x * chains_per_block + local_chain;
    if (chain >= a.chain_end) return;

    double *cs = smem + model_smem_doubles(sm.family, D, DP, sm.K) + (size_t) local_chain * chain_smem_doubles(DP, G);
    double *scratch = cs;
    double *pre_u = cs + DP;                                    // [G][kPre]
    uint32_t *pre_k = (uint32_t *) (cs + DP + (size_t) G * kPre);  // [G][4]: after_key, run_key

    const double contour = *a.contour;
    const int S = a.S, kph = a.k;

    // ---- chain prelude (bases.py:64; uni_slice_sampler.py:343-358, :410-413)
    const Key chain_key = split_child(a.key, (uint64_t) chain);
    const Key sample_key = split_child(chain_key, 0);
    const Key seed_key = split_child(chain_key, 1);
    const double useed = uniform01(seed_key, 0);
    const long long sidx = seed_index(a.live_logL, a.seed_table, a.N, contour, useed);
    double U0[DPL], d[DPL], x[DPL];
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        U0[s] = (j < D) ? a.live_U[sidx * D + j] : 0.5;
    }
    double logL0 = a.live_logL[sidx];
    const Key direction_key = split_child(sample_key, 0);
    const Key sample_key2 = split_child(sample_key, 1);
    sample_direction<DPL>(g, D, direction_key, d);

    const long long out_row = chain - a.chain_begin;
    long long nev = 0;

    for (int base = 0; base < S; base += G) {
        // ---- precompute the key stream of slices base .. base+G-1, one slice per lane
        {
            const int j = base + g.lane;
            if (j < S) {
                const Key slice_key = split_child(sample_key2, (uint64_t) j);  // :420
                Key run_key = split_child(slice_key, 0);                       // :201 (child 1 = n_key unused)
                Key t_key = split_child(slice_key, 2);
                const Key after_key = split_child(slice_key, 3);
                pre_u[g.lane * kPre + 0] = uniform01(t_key, 0);
#pragma unroll
                for (int p = 1; p < kPre; ++p) {
                    t_key = split_child(run_key, 1);  // :169 (child 2 = shrink_key unused)
                    run_key = split_child(run_key, 0);
                    pre_u[g.lane * kPre + p] = uniform01(t_key, 0);
                }
                uint32_t *pk = pre_k + g.lane * 4;
                pk[0] = after_key.a;
                pk[1] = after_key.b;
                pk[2] = run_key.a;
                pk[3] = run_key.b;
            }
        }
        group_sync(g);
        const int jmax = min(G, S - base);
        for (int jl = 0; jl < jmax; ++jl) {
            const int j = base + jl;
            const double alpha = alpha_schedule(j, S);
            const double *uq = pre_u + jl * kPre;
            const uint32_t *pk = pre_k + jl * 4;
            double left, right;
            slice_bounds<DPL>(g, D, U0, d, left, right);
            double t = left + uq[0] * (right - left);  // _pick_point_in_interval :83-85
#pragma unroll
            for (int s = 0; s < DPL; ++s) x[s] = fma(t, d[s], U0[s]);
            double logL = forward_group<DPL>(sm, g, x, scratch);
            int ne = 1;
            Key run_key = Key{pk[2], pk[3]};
            // shrink loop (:160-196)
            while (!((logL > contour) || ((logL0 == contour) && (logL == contour)))) {
                if (t < 0.0) left = t;  // _shrink_interval :92-111
                if (t > 0.0) right = t;
                if (a.midpoint) {
                    if (t < 0.0) left = alpha * left;
                    if (t > 0.0) right = alpha * right;
                }
                double uu;
                if (ne < kPre) {
                    uu = uq[ne];
                } else {
                    const Key t_key = split_child(run_key, 1);
                    run_key = split_child(run_key, 0);
                    uu = uniform01(t_key, 0);
                }
                t = left + uu * (right - left);
#pragma unroll
                for (int s = 0; s < DPL; ++s) x[s] = fma(t, d[s], U0[s]);
                logL = forward_group<DPL>(sm, g, x, scratch);
                ne += 1;
            }
#pragma unroll
            for (int s = 0; s < DPL; ++s) U0[s] = x[s];
            logL0 = logL;
            nev += ne;
            sample_direction<DPL>(g, D, Key{pk[0], pk[1]}, d);  // :272
            // phantom capture: cumulative_samples[-(k+1):-1] (:430-440)
            if (kph > 0 && j >= S - 1 - kph && j < S - 1) {
                const long long slot = out_row * kph + (j - (S - 1 - kph));
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) {
                        if (a.ph_U) a.ph_U[slot * D + jj] = U0[s];
                        if (a.packed)
                            a.packed[out_row * a.packed_row_doubles + (D + 2) + (long long) (j - (S - 1 - kph)) * (D + 1) + jj] = U0[s];
                    }
                }
                if (g.lane == 0) {
                    if (a.ph_logL) a.ph_logL[slot] = logL0;
                    if (a.packed)
                        a.packed[out_row * a.packed_row_doubles + (D + 2) + (long long) (j - (S - 1 - kph)) * (D + 1) + D] = logL0;
                }
            }
        }
        group_sync(g);  // pre_u / pre_k are rewritten by the next chunk
    }
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D) {
            if (a.out_U) a.out_U[out_row * D + j] = U0[s];
            if (a.packed) a.packed[out_row * a.packed_row_doubles + j] = U0[s];
        }
    }
    if (g.lane == 0) {
        if (a.out_logL) a.out_logL[out_row] = logL0;
        if (a.out_nevals) a.out_nevals[out_row] = nev;
        if (a.packed) {
            a.packed[out_row * a.packed_row_doubles + D] = logL0;
            a.packed[out_row * a.packed_row_doubles + D + 1] = __longlong_as_double(nev);
        }
    }
}

template <int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_slice_chains(SliceArgs a) {
    extern __shared__ double smem[];
    slice_chains_body<DPL>(a, smem);
}

// ---- prior draws for the initial live set / uniform rejection sampler --------------------------
struct DrawArgs {
    NsModelDesc model;
    Key key;                // keys = split(key, n_total)
    const double *contour;  // device scalar or nullptr (init: -inf)
    double *out_U;
    double *out_logL;
    long long *out_nevals;
    long long begin, end;
    int G;
    int uniform_sampler;  // 0: _single_uniform_sample ; 1: UniformSampler._get_sample
};

// Model.sample_U: uniform(split(key, 2)[1], (D,))
template <int DPL>
__device__ __forceinline__ void sample_U(const Grp &g, int D, Key key, double (&u)[DPL]) {
    const Key k = split_child(key, 1);
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * g.G + g.lane;
        u[s] = (j < D) ? uniform01(k, (uint64_t) j) : 0.5;
    }
}

template <int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_draw(DrawArgs a) {
    extern __shared__ double smem[];
    const int G = a.G, DP = G * DPL, D = a.model.D;
    ModelSmem sm;
    stage_model(a.model, DP, smem, sm);
    __syncthreads();
    const Grp g = make_group(G);
    const int per_block = kThreadsPerBlock / G;
    const int local = threadIdx.x / G;
    const long long i = a.begin + (long long) blockIdx.x * per_block + local;
    if (i >= a.end) return;
    double *scratch = smem + model_smem_doubles(sm.family, D, DP, sm.K) + (size_t) local * DP;
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    const double contour = a.contour ? *a.contour : -kInf;
    const Key k0 = split_child(a.key, (uint64_t) i);
    Key key = split_child(k0, 0);
    Key sk = split_child(k0, 1);
    double u[DPL];
    sample_U<DPL>(g, D, sk, u);
    double logL = forward_group<DPL>(sm, g, u, scratch);
    long long ne = 1;
    for (;;) {
        bool done;
        if (a.uniform_sampler) done = (logL > contour) || (logL == contour) || (ne >= 100);
        else done = !(logL <= contour);
        if (done) break;
        sk = split_child(key, 1);
        key = split_child(key, 0);
        sample_U<DPL>(g, D, sk, u);
        logL = forward_group<DPL>(sm, g, u, scratch);
        ne += 1;
    }
    const long long o = i - a.begin;
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D) a.out_U[o * D + j] = u[s];
    }
    if (g.lane == 0) {
        a.out_logL[o] = logL;
        a.out_nevals[o] = ne;
    }
}

// ---- vmap(Model.forward) / vmap(Model.transform) -----------------------------------------------
struct ForwardArgs {
    NsModelDesc model;
    const double *U;
    double *out_logL;
    double *out_X;
    long long n;
    int G;
};

template <int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_forward(ForwardArgs a) {
    extern __shared__ double smem[];
    const int G = a.G, DP = G * DPL, D = a.model.D;
    ModelSmem sm;
    stage_model(a.model, DP, smem, sm);
    __syncthreads();
    const Grp g = make_group(G);
    const int per_block = kThreadsPerBlock / G;
    const int local = threadIdx.x / G;
    const long long i = (long long) blockIdx.x * per_block + local;
    if (i >= a.n) return;
    double *scratch = smem + model_smem_doubles(sm.family, D, DP, sm.K) + (size_t) local * DP;
    double u[DPL], X[DPL];
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        u[s] = (j < D) ? a.U[i * D + j] : 0.5;
    }
    transform_dims<DPL>(sm, g, u, X);
    const double logL = loglik_group<DPL>(sm, g, X, scratch);
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D && a.out_X) a.out_X[i * D + j] = X[s];
    }
    if (g.lane == 0 && a.out_logL) a.out_logL[i] = logL;
}

}  // namespace nsb

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
//
// Two facts about the reference's slice move make it GPU-friendly without changing its results:
//  (1) the key stream of a slice (run_key chain, t_key draws, after_key) does not depend on any
//      likelihood value, so lane L of a group precomputes the stream of slice base+L while its
//      neighbours do the same for theirs, and the direction of the next slice can be drawn while the
//      current one is still being evaluated;
//  (2) a rejected proposal shrinks the bracket to its own t (times alpha), which is known before
//      its likelihood is: the next P proposals of the shrink loop can be generated speculatively and
//      evaluated together; the first accepted one wins and n_evals counts only up to it, exactly as
//      the sequential loop would.
#pragma once
#include "ns_model.cuh"
#include "ns_types.cuh"

namespace nsb {

#ifdef NSB_PROFILE
// cycle accounting of chain 0's critical path (debug builds only): g_prof[k] accumulates clock64 deltas
__device__ unsigned long long g_prof[16];
#define NSB_T0() long long _t0 = clock64()
#define NSB_TICK(k)                               \
    {                                             \
        const long long _t1 = clock64();          \
        prof[k] += (unsigned long long) (_t1 - _t0); \
        _t0 = _t1;                                \
    }
#else
#define NSB_T0()
#define NSB_TICK(k)
#endif

#ifdef NSB_TIMELINE
// globaltimer stamps of the generator / slice kernel pair of the latest body (debug builds only):
// [0] first generator CTA start, [1] last generator CTA end, [2] first slice CTA start, [3] last slice CTA end,
// [4..7] the same accumulated relative to [0] over all bodies, [8] bodies
__device__ unsigned long long g_tl[16];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

// Uniforms per slice drawn ahead by the lane-parallel precompute (shrink steps beyond that fall
// back to walking the run_key chain inline).
#ifndef NSB_KPRE
#define NSB_KPRE 8
#endif
constexpr int kPre = NSB_KPRE;

// Watchdog of the shrink loop: a bracket that has collapsed onto the seed point re-evaluates the seed
// itself, which satisfies the constraint, so a deterministic likelihood accepts within ~2200 halvings of a double.
// A chain still shrinking after this many proposals of ONE slice faces a non-deterministic (or NaN at the seed)
// likelihood: it stays where it is and NSB200_ERR_SHRINK_LOOP is raised instead of hanging the GPU.
constexpr int kMaxShrinkProposals = 1 << 16;

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
    int S, k, midpoint;
    // optional packed-row output for the multi-GPU all-gather (row = [U[D], logL, nevals, k x (U[D], logL)])
    double *packed;
    long long packed_row_doubles;
    // engine mode: key, contour and the current live buffer come from the device-resident control
    // block, so the host never has to know them (and the launch is a no-op once the loop is done)
    const DevCtl *ctl;
    LiveSet live0, live1;
    // optional precomputed per-chain streams (k_chain_streams): directions [n][S][D], proposal
    // uniforms [n][S][kPre] and the run_key after kPre-1 shrink steps [n][S]; n = chain_end - chain_begin
    const double *pre_dirs;
    const double *pre_us;
    const uint2 *pre_rkeys;
    const double *alpha_tab;  // optional [S]: alpha_schedule(j, S) (saves an int->double division per slice)
    // fused all-gather (multi-GPU engines connected over CUDA IPC): every packed row is also stored straight
    // into the gather buffer of each rank (own buffer included) through NVLink peer mappings, so no collective
    // kernel runs between the chains and the merge -- only an arrival barrier (k_peer_barrier)
    int n_peers;
    double *peers[8];  // this rank's block inside rank r's gather buffer
    int *err;          // optional device word: NSB200_ERR_* bits are OR-ed in (shrink-loop watchdog)
};

__device__ __forceinline__ void packed_store(const SliceArgs &a, long long off, double v) {
    if (a.packed) a.packed[off] = v;
    for (int r = 0; r < a.n_peers; ++r) a.peers[r][off] = v;
}

// jnp.linspace(0.5, 1., S)[j]
__device__ __forceinline__ double alpha_schedule(int j, int S) {
    if (S == 1) return 0.5;
    const int div = S - 1;
    if (j == div) return 1.0;
    const double step = (double) j / (double) div;
    return 0.5 * (1.0 - step) + 1.0 * step;
}

// _sample_direction: d = normal(key, (D,)); d /= ||d||
// direction / its norm (_sample_direction :23-38): the reciprocal-Newton quotient (within 1 ulp of the IEEE division,
// 8 instructions instead of ~20) -- one per deviate of the stream generator.  -DNSB_FAST_DIR_DIV=0: IEEE division.
#ifndef NSB_FAST_DIR_DIV
#define NSB_FAST_DIR_DIV 1
#endif
#if NSB_FAST_DIR_DIV
#define NSB_DIR_DIV(a, b) fast_div((a), (b))
#else
#define NSB_DIR_DIV(a, b) ((a) / (b))
#endif
template <int G, int DPL>
__device__ __forceinline__ void sample_direction(const Grp<G> &g, int D, Key key, double (&d)[DPL]) {
    if (D == 1) {
#pragma unroll
        for (int s = 0; s < DPL; ++s) d[s] = (s == 0 && g.lane == 0) ? 1.0 : 0.0;
        return;
    }
    double ss = 0.0;
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        d[s] = (j < D) ? normal_from_bits(bits64(key, (uint64_t) j)) : 0.0;
        ss = fma(d[s], d[s], ss);
    }
    const double nrm = sqrt(group_sum(g, ss));
#pragma unroll
    for (int s = 0; s < DPL; ++s) d[s] = NSB_DIR_DIV(d[s], nrm);
}

// Exact min over the lanes in `mask` of NON-NEGATIVE doubles (+inf allowed): for such values the IEEE
// bit pattern orders like an unsigned integer, so two REDUX.MIN instructions (high word, then low
// word among the lanes that hold the winning high word) replace a 5-step 64-bit shuffle butterfly.
__device__ __forceinline__ double group_min_nonneg(unsigned mask, double v) {
    const unsigned hi = (unsigned) __double2hiint(v), lo = (unsigned) __double2loint(v);
    const unsigned mhi = __reduce_min_sync(mask, hi);
    const unsigned mlo = __reduce_min_sync(mask, hi == mhi ? lo : 0xFFFFFFFFu);
    return __hiloint2double((int) mhi, (int) mlo);
}

// _slice_bounds: intersection of the line U0 + t d with the unit cube.
template <int G, int DPL>
__device__ __forceinline__ void slice_bounds(const Grp<G> &g, int D, const double (&U0)[DPL], const double (&d)[DPL],
                                             double &left, double &right) {
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    double r = kInf, nl = kInf;  // right bound and MINUS the left bound, both >= 0
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D) {
            const double t1 = fast_div(1.0 - U0[s], d[s]);
            const double t0 = fast_div(-U0[s], d[s]);
            if (t1 >= 0.0) r = fmin(r, t1);
            if (t1 <= 0.0) nl = fmin(nl, -t1);
            if (t0 >= 0.0) r = fmin(r, t0);
            if (t0 <= 0.0) nl = fmin(nl, -t0);
        }
    }
    if (G > 1) {
        // -0.0 would order above every positive number as an integer: canonicalise the zeros
        r = group_min_nonneg(g.m(), r + 0.0);
        nl = group_min_nonneg(g.m(), nl + 0.0);
    }
    right = r;
    left = -nl;
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
//   [model image][per chain: scratch[DP][P] | pre_u[G][kPre] | pre_keys[G][4 x u32]]
__host__ __device__ inline size_t chain_smem_doubles(int G, int DPL, int P, bool slice) {
    return (size_t) G * DPL * P + (slice ? (size_t) G * (kPre + 2) : 0);
}

// W > 1 ("warp team", experiment, NOT instantiated: measured slower on B200 -- register spills at 2 warps x
// 1600 chains -- kept as the starting point of a producer/consumer split): the CTA is ONE chain run by W warps that execute the same program; in every
// shrink round warp w evaluates the w-th speculative proposal and the W log-likelihoods are exchanged
// through shared memory around one __syncthreads().  Unlike the in-warp batch (P) this adds real
// parallelism: a config-2 sized problem leaves most warp slots and issue cycles idle, and the
// kernel time is the latency of its longest chain, so halving the sequential rounds per slice is
// worth more than the ~15 % of evaluations that are thrown away.  Results are unchanged (first
// accepted proposal wins, n_evals counts up to it).  Requires G == 32, P == 1, PRE.
template <int G, int DPL, int P, int FAM, bool PRE, int W = 1, bool RS = false>
__device__ __forceinline__ void slice_chains_body(const SliceArgs &a, double *smem) {
    static_assert(W == 1 || (G == 32 && P == 1 && PRE), "warp teams need G == 32, P == 1 and precomputed streams");
    constexpr bool TEAM = W > 1;
    constexpr int DP = G * DPL;
    const int D = a.model.D;
    Key base_key = a.key;
    const double *contour_ptr = a.contour;
    const double *live_U = a.live_U;
    const double *live_logL = a.live_logL;
    if (a.ctl) {
        if (!a.ctl->active) return;
        const LiveSet &live = a.ctl->cur ? a.live1 : a.live0;
        base_key = a.ctl->sample_key;
        contour_ptr = &a.ctl->contour;
        live_U = live.U;
        live_logL = live.logL;
    }
    ModelSmem sm;
    stage_model<G, DPL, RS>(a.model, smem, sm);
    __syncthreads();
    const Grp<G> g;
    const int wt = TEAM ? (int) (threadIdx.x >> 5) : 0;  // warp of the team = which proposal of a round it evaluates
    const int chains_per_block = TEAM ? 1 : blockDim.x / G;  // the slice kernel picks its CTA size at launch
    const int local_chain = TEAM ? wt : threadIdx.x / G;     // index of this lane group's private smem slot
    const long long chain = a.chain_begin + (long long) blockIdx.x * chains_per_block + (TEAM ? 0 : local_chain);
    if (chain >= a.chain_end) return;
    double *res = smem + model_smem_doubles(a.model.family, D, G, DPL, a.model.K, RS) +
                  (size_t) (TEAM ? W : 0) * chain_smem_doubles(G, DPL, P, true);  // [2][W] exchange buffer (TEAM)
    int par = 0;
    DenseRow<G, DPL, RS> row;
    row.load(sm, g.lane);

    double *cs = smem + model_smem_doubles(sm.family, D, G, DPL, sm.K, RS) +
                 (size_t) local_chain * chain_smem_doubles(G, DPL, P, true);
    double *scratch = cs;
    double *pre_u = cs + DP * P;                                       // [G][kPre]
    uint32_t *pre_k = (uint32_t *) (cs + DP * P + (size_t) G * kPre);  // [G][4]: after_key, run_key

    const double contour = *contour_ptr;
    const int S = a.S, kph = a.k;
    const bool midpoint = a.midpoint != 0;

    // ---- chain prelude (bases.py:64; uni_slice_sampler.py:343-358, :410-413)
    const Key chain_key = split_child(base_key, (uint64_t) chain);
    const Key sample_key = split_child(chain_key, 0);
    const Key seed_key = split_child(chain_key, 1);
    const double useed = uniform01(seed_key, 0);
    const long long sidx = seed_index(live_logL, a.seed_table, a.N, contour, useed);
    double U0[DPL], d[DPL];
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        U0[s] = (j < D) ? live_U[sidx * D + j] : 0.5;
    }
    double logL0 = live_logL[sidx];
    const long long out_row = chain - a.chain_begin;
    long long nev = 0;
#ifdef NSB_PROFILE
    unsigned long long prof[16];
    for (int k = 0; k < 16; ++k) prof[k] = 0;
#endif
    NSB_T0();
    Key sample_key2 = Key{0, 0};
    // stream of slice 0 (PRE: fetched; otherwise derived below)
    double unext = 0.0;           // lane p < kPre holds uniform p of the next slice
    uint2 rknext = make_uint2(0, 0);
    if (PRE) {
        const long long sb = out_row * S;
#pragma unroll
        for (int s = 0; s < DPL; ++s) {
            const int j = s * G + g.lane;
            d[s] = (j < D) ? __ldg(a.pre_dirs + sb * D + j) : 0.0;
        }
        if (g.lane < kPre) unext = __ldg(a.pre_us + sb * kPre + g.lane);
        rknext = __ldg(a.pre_rkeys + sb);
    } else {
        const Key direction_key = split_child(sample_key, 0);
        sample_key2 = split_child(sample_key, 1);
        sample_direction<G, DPL>(g, D, direction_key, d);
    }

    NSB_TICK(0)  // chain prelude
    for (int base = 0; base < S; base += G) {
        // ---- precompute the key stream of slices base .. base+G-1, one slice per lane
        if (!PRE) {
            const int j = base + g.lane;
            if (j < S) {
                const Key slice_key = split_child(sample_key2, (uint64_t) j);  // :420
                Key run_key = split_child(slice_key, 0);                       // :201 (child 1 = n_key unused)
                Key t_key = split_child(slice_key, 2);
                const Key after_key = split_child(slice_key, 3);
                pre_u[g.lane * kPre + 0] = uniform01(t_key, 0);
#pragma unroll 1
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
            const double alpha = a.alpha_tab ? __ldg(a.alpha_tab + j) : alpha_schedule(j, S);
            const double *uq;
            Key run_key;
            double dnext[DPL];
            if (PRE) {
                // this slice's uniforms were fetched one slice ago: park them in shared memory (the
                // shrink loop indexes them), then issue the loads of the next slice's stream so that
                // their DRAM/L2 latency hides behind this slice's evaluations
                if (g.lane < kPre) pre_u[g.lane] = unext;
                run_key = Key{rknext.x, rknext.y};
                uq = pre_u;
                const long long sn = out_row * S + j + 1;
                const bool more = j + 1 < S;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    dnext[s] = (more && jj < D) ? __ldg(a.pre_dirs + sn * D + jj) : 0.0;
                }
                if (more && g.lane < kPre) unext = __ldg(a.pre_us + sn * kPre + g.lane);
                if (more) rknext = __ldg(a.pre_rkeys + sn);
                group_sync(g);
            } else {
                uq = pre_u + jl * kPre;
                const uint32_t *pk = pre_k + jl * 4;
                // direction of the NEXT slice (:272): independent of this slice's evaluations
                sample_direction<G, DPL>(g, D, Key{pk[0], pk[1]}, dnext);
                run_key = Key{pk[2], pk[3]};
            }
            NSB_TICK(1)  // stream fetch / direction
            double left, right;
            slice_bounds<G, DPL>(g, D, U0, d, left, right);
            NSB_TICK(2)  // bounds
            int ne = 0;  // proposals generated so far in this slice
            double logL_acc = 0.0;
            double u_ahead = uq[0];
#ifndef NSB_U_AHEAD
#define NSB_U_AHEAD 0
#endif
            // P > 1 (-DNSB_U_AHEAD=1): the round's P uniforms are loaded one round ahead, like u_ahead for P = 1
            double u_pre[P];
            if (NSB_U_AHEAD && P > 1) {
#pragma unroll
                for (int p = 0; p < P; ++p) u_pre[p] = uq[p < kPre ? p : kPre - 1];
            }
            if (TEAM) {
                for (;;) {
                    double ts[W];
                    double l = left, r = right;
#pragma unroll
                    for (int p = 0; p < W; ++p) {
                        double uu;
                        const int n = ne + p;
                        if (n < kPre) {
                            uu = uq[n];
                        } else {
                            const Key t_key = split_child(run_key, 1);
                            run_key = split_child(run_key, 0);
                            uu = uniform01(t_key, 0);
                        }
                        const double t = l + uu * (r - l);
                        ts[p] = t;
                        if (t < 0.0) l = midpoint ? alpha * t : t;
                        if (t > 0.0) r = midpoint ? alpha * t : t;
                    }
                    double tmine = ts[0];
#pragma unroll
                    for (int p = 1; p < W; ++p) tmine = (wt == p) ? ts[p] : tmine;
                    double x1[1][DPL], l1[1];
#pragma unroll
                    for (int s = 0; s < DPL; ++s) x1[0][s] = fma(tmine, d[s], U0[s]);
                    forward_group<G, DPL, 1, FAM, RS>(sm, g, row, x1, scratch, l1);
                    if (g.lane == 0) res[par * W + wt] = l1[0];
                    __syncthreads();
                    int hit = -1;
                    double l_hit = 0.0, t_hit = 0.0;
#pragma unroll
                    for (int p = W - 1; p >= 0; --p) {
                        const double lp = res[par * W + p];
                        if ((lp > contour) || ((logL0 == contour) && (lp == contour))) {
                            hit = p;
                            l_hit = lp;
                            t_hit = ts[p];
                        }
                    }
                    par ^= 1;
                    if (hit >= 0) {
#pragma unroll
                        for (int s = 0; s < DPL; ++s) U0[s] = fma(t_hit, d[s], U0[s]);
                        logL_acc = l_hit;
                        ne += hit + 1;
                        break;
                    }
                    ne += W;
                    left = l;
                    right = r;
                    if (ne >= kMaxShrinkProposals) {  // watchdog: stay at the current point, flag the run
                        if (a.err && g.lane == 0 && wt == 0) atomicOr(a.err, NSB200_ERR_SHRINK_LOOP);
                        logL_acc = logL0;
                        break;
                    }
                }
            } else
            for (;;) {
                // ---- generate P proposals assuming each previous one is rejected (:92-111, :169-186)
                double ts[P], x[P][DPL], logL[P];
                double l = left, r = right;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    double uu;
                    const int n = ne + p;
                    if (P == 1 && n < kPre) {
                        uu = u_ahead;  // loaded while the previous proposal was being evaluated
                    } else if (n < kPre) {
                        uu = (NSB_U_AHEAD && P > 1) ? u_pre[p] : uq[n];
                    } else {
                        const Key t_key = split_child(run_key, 1);
                        run_key = split_child(run_key, 0);
                        uu = uniform01(t_key, 0);
                    }
                    const double t = l + uu * (r - l);  // _pick_point_in_interval :83-85
                    ts[p] = t;
                    if (t < 0.0) l = midpoint ? alpha * t : t;  // _shrink_interval
                    if (t > 0.0) r = midpoint ? alpha * t : t;
#pragma unroll
                    for (int s = 0; s < DPL; ++s) x[p][s] = fma(t, d[s], U0[s]);
                }
                if (P == 1) u_ahead = uq[min(ne + 1, kPre - 1)];  // next proposal's uniform: hide the smem latency
                if (NSB_U_AHEAD && P > 1) {
#pragma unroll
                    for (int p = 0; p < P; ++p) u_pre[p] = uq[min(ne + P + p, kPre - 1)];
                }
                NSB_TICK(3)  // proposal generation
#ifdef NSB_PROFILE
                {
                    double Xp[P][DPL];
                    transform_dims<G, DPL, P>(sm, g, x, Xp);
                    NSB_TICK(4)  // prior transform
                    loglik_group<G, DPL, P, FAM, RS>(sm, g, row, Xp, scratch, logL);
                    NSB_TICK(5)  // likelihood
                    prof[8] += 1;
                    {
                        bool tl = false, tl2 = false;
                        for (int s = 0; s < DPL; ++s) {
                            tl |= !(-log(4.0 * (x[0][s] * (1.0 - x[0][s]))) < 6.25);
                            tl2 |= !(fabs(x[0][s] - 0.5) <= 0.425);
                        }
                        prof[10] += __any_sync(g.m(), tl) ? 1 : 0;
                        prof[11] += __any_sync(g.m(), tl2) ? 1 : 0;
                    }
                }
#else
                forward_group<G, DPL, P, FAM, RS>(sm, g, row, x, scratch, logL);
#endif
                // ---- first accepted proposal wins (:160-166)
                int hit = -1;
#pragma unroll
                for (int p = P - 1; p >= 0; --p) {
                    const bool ok = (logL[p] > contour) || ((logL0 == contour) && (logL[p] == contour));
                    if (ok) hit = p;
                }
                if (hit >= 0) {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        if (p == hit) {
#pragma unroll
                            for (int s = 0; s < DPL; ++s) U0[s] = x[p][s];
                            logL_acc = logL[p];
                        }
                    }
                    ne += hit + 1;
                    break;
                }
                ne += P;
                left = l;
                right = r;
                if (ne >= kMaxShrinkProposals) {  // watchdog: stay at the current point, flag the run
                    if (a.err && g.lane == 0) atomicOr(a.err, NSB200_ERR_SHRINK_LOOP);
                    logL_acc = logL0;
                    break;
                }
            }
            NSB_TICK(6)  // accept logic / loop overhead
            logL0 = logL_acc;
            nev += ne;
#ifdef NSB_PROFILE
            prof[9] += 1;
#endif
#pragma unroll
            for (int s = 0; s < DPL; ++s) d[s] = dnext[s];
            // phantom capture: cumulative_samples[-(k+1):-1] (:430-440)
            if (kph > 0 && j >= S - 1 - kph && j < S - 1 && wt == 0) {
                const long long slot = out_row * kph + (j - (S - 1 - kph));
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) {
                        if (a.ph_U) a.ph_U[slot * D + jj] = U0[s];
                        packed_store(a, out_row * a.packed_row_doubles + (D + 2) + (long long) (j - (S - 1 - kph)) * (D + 1) + jj, U0[s]);
                    }
                }
                if (g.lane == 0) {
                    if (a.ph_logL) a.ph_logL[slot] = logL0;
                    packed_store(a, out_row * a.packed_row_doubles + (D + 2) + (long long) (j - (S - 1 - kph)) * (D + 1) + D, logL0);
                }
            }
        }
        group_sync(g);  // pre_u / pre_k are rewritten by the next chunk
    }
    NSB_TICK(7)
#ifdef NSB_PROFILE
    if (chain == a.chain_begin && g.lane == 0 && wt == 0)
        for (int k = 0; k < 16; ++k) atomicAdd(&g_prof[k], prof[k]);
#endif
    if (wt != 0) return;
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D) {
            if (a.out_U) a.out_U[out_row * D + j] = U0[s];
            packed_store(a, out_row * a.packed_row_doubles + j, U0[s]);
        }
    }
    if (g.lane == 0) {
        if (a.out_logL) a.out_logL[out_row] = logL0;
        if (a.out_nevals) a.out_nevals[out_row] = nev;
        packed_store(a, out_row * a.packed_row_doubles + D, logL0);
        packed_store(a, out_row * a.packed_row_doubles + D + 1, __longlong_as_double(nev));
    }
}

// -DNSB_SLICE_MIN_BLOCKS=n (experiments): resident CTAs per SM the compiler must allow for, i.e. a register cap of
// 65536 / (n * kThreadsPerBlock) per thread (kThreadsPerBlock = 128: n = 4 -> 128 registers).
#ifdef NSB_SLICE_MIN_BLOCKS
#define NSB_SLICE_BOUNDS __launch_bounds__(kThreadsPerBlock, NSB_SLICE_MIN_BLOCKS)
#else
#define NSB_SLICE_BOUNDS __launch_bounds__(kThreadsPerBlock)
#endif
template <int G, int DPL, int P>
__global__ void NSB_SLICE_BOUNDS k_slice_chains(SliceArgs a) {
    extern __shared__ double smem[];
#ifdef NSB_TIMELINE
    if (threadIdx.x == 0) atomicMin(&g_tl[2], gtime());
    struct Stamp { __device__ ~Stamp() { if (threadIdx.x == 0) atomicMax(&g_tl[3], gtime()); } } stamp;
#endif
    if (a.pre_dirs) {
        NSB_FAMILY_SWITCH(a.model.family, slice_chains_body<G, DPL, P, kFam, true>(a, smem));
    } else {
        NSB_FAMILY_SWITCH(a.model.family, slice_chains_body<G, DPL, P, kFam, false>(a, smem));
    }
}

// Warp team: one CTA = one chain run by W warps; in every shrink round warp w evaluates the w-th speculative
// proposal (see slice_chains_body).  Dense Gaussian, 17 <= D <= 32, pre-generated streams.  The factor is staged in
// shared memory (RS) so that the chain state fits 88 registers: W x 1600 chains have to be resident in ONE wave
// (11 CTAs of 64 threads per SM), otherwise the second wave costs more than the halved rounds save.
template <int W>
__global__ void __launch_bounds__(32 * W, W == 2 ? 11 : 5) k_slice_chains_team(SliceArgs a) {
    extern __shared__ double smem[];
    slice_chains_body<32, 1, 1, NSB200_FAM_GAUSS_DENSE, true, W, true>(a, smem);
}

__global__ void k_alpha_table(int S, double *out) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < S; j += gridDim.x * blockDim.x) out[j] = alpha_schedule(j, S);
}

// ---- data-independent per-chain streams ----------------------------------------------------------
// Nothing in a chain's key tree depends on a likelihood value, so the directions of all S slices,
// the first kPre proposal uniforms of every slice and the run_key to continue from can be produced
// by a throughput-bound kernel (one warp per chain x 32 slices) instead of on the chain's critical
// path.  Layouts match what the slice kernel reads: dirs [n][S][D], us [n][S][kPre], rkeys [n][S].
struct StreamArgs {
    Key key;
    const DevCtl *ctl;  // engine mode: key = ctl->stream_key[key_slot]
    int key_slot;
    long long chain_begin, chain_end;
    int S, D;
    double *dirs;
    double *us;
    uint2 *rkeys;
};

// QMAX = register slots per direction: dimensions lane + 32 q, q < QMAX (D <= 32 QMAX)
template <int QMAX>
__global__ void __launch_bounds__(1024) k_chain_streams(StreamArgs a) {
    // Programmatic dependent launch: this CTA is resident, so the slice kernel queued behind the generator may
    // start now and take the SMs the generator's CTAs did not claim (no-op for a normally launched successor).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef NSB_TIMELINE
    if (threadIdx.x == 0) atomicMin(&g_tl[0], gtime());
    struct Stamp { __device__ ~Stamp() { if (threadIdx.x == 0) atomicMax(&g_tl[1], gtime()); } } stamp;
#endif
    Key base_key = a.key;
    if (a.ctl) base_key = a.ctl->stream_key[a.key_slot];
    const int lane = threadIdx.x & 31;
    const int S = a.S, D = a.D;
    const int n_chunks = (S + 31) / 32;
    // persistent warps over (chain, 32-slice chunk) items: the grid is kept small on purpose (a few
    // warps per SM) so that this throughput-bound kernel back-fills issue slots next to the
    // latency-bound slice kernel instead of starving its warps
    const long long n_items = (a.chain_end - a.chain_begin) * n_chunks;
    const long long n_warps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long wid = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < n_items; wid += n_warps) {
    const long long row = wid / n_chunks;
    const int base = (int) (wid - row * n_chunks) * 32;
    const long long chain = a.chain_begin + row;
    const Key chain_key = split_child(base_key, (uint64_t) chain);
    const Key sample_key = split_child(chain_key, 0);  // bases.py:64
    const Key sample_key2 = split_child(sample_key, 1);  // uni_slice_sampler.py:410
    // phase 1: lane L owns slice base + L
    Key after_key = Key{0, 0};
    {
        const int j = base + lane;
        if (j < S) {
            const Key slice_key = split_child(sample_key2, (uint64_t) j);  // :420
            Key run_key = split_child(slice_key, 0);                       // :201
            Key t_key = split_child(slice_key, 2);
            after_key = split_child(slice_key, 3);
            double *u = a.us + (row * S + j) * kPre;
            u[0] = uniform01(t_key, 0);
#pragma unroll 1
            for (int p = 1; p < kPre; ++p) {
                t_key = split_child(run_key, 1);  // :169
                run_key = split_child(run_key, 0);
                u[p] = uniform01(t_key, 0);
            }
            a.rkeys[row * S + j] = make_uint2(run_key.a, run_key.b);
        }
    }
    // phase 2: lanes = dimensions; direction of slice j+1 comes from after_key_j (:272), the first
    // one from direction_key (:410-413).  (Two directions per trip with the Threefry block inlined was measured: no
    // faster -- 32 warps per SM already fill the issue slots -- and the run 2 % slower.)
    const int jl0 = (base == 0) ? -1 : 0;
#pragma unroll 1
    for (int jl = jl0; jl < 32; ++jl) {
        const int jdst = base + jl + 1;  // slice that uses this direction
        if (jdst >= S) break;
        Key k;
        if (jl < 0) {
            k = split_child(sample_key, 0);
        } else {
            k.a = __shfl_sync(0xFFFFFFFFu, after_key.a, jl);
            k.b = __shfl_sync(0xFFFFFFFFu, after_key.b, jl);
        }
        double *dst = a.dirs + (row * S + jdst) * D;
        if (D == 1) {
            if (lane == 0) dst[0] = 1.0;
            continue;
        }
        double v[QMAX];
        double ss = 0.0;
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int j = lane + 32 * q;
            v[q] = (j < D) ? normal_from_bits(bits64(k, (uint64_t) j)) : 0.0;
            ss = fma(v[q], v[q], ss);
            if (32 * (q + 1) >= D) break;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        const double nrm = sqrt(ss);
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int j = lane + 32 * q;
            if (j < D) dst[j] = NSB_DIR_DIV(v[q], nrm);
            if (32 * (q + 1) >= D) break;
        }
    }
    }
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
    int uniform_sampler;  // 0: _single_uniform_sample ; 1: UniformSampler._get_sample
};

// Model.sample_U: uniform(split(key, 2)[1], (D,))
template <int G, int DPL>
__device__ __forceinline__ void sample_U(const Grp<G> &g, int D, Key key, double (&u)[1][DPL]) {
    const Key k = split_child(key, 1);
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        u[0][s] = (j < D) ? uniform01(k, (uint64_t) j) : 0.5;
    }
}

template <int G, int DPL, int FAM>
__device__ __forceinline__ void draw_body(const DrawArgs &a, double *smem) {
    constexpr int DP = G * DPL;
    const int D = a.model.D;
    ModelSmem sm;
    stage_model<G, DPL>(a.model, smem, sm);
    __syncthreads();
    const Grp<G> g;
    constexpr int per_block = kThreadsPerBlock / G;
    const int local = threadIdx.x / G;
    const long long i = a.begin + (long long) blockIdx.x * per_block + local;
    if (i >= a.end) return;
    DenseRow<G, DPL> row;
    row.load(sm, g.lane);
    double *scratch = smem + model_smem_doubles(sm.family, D, G, DPL, sm.K) + (size_t) local * DP;
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    const double contour = a.contour ? *a.contour : -kInf;
    const Key k0 = split_child(a.key, (uint64_t) i);
    Key key = split_child(k0, 0);
    Key sk = split_child(k0, 1);
    double u[1][DPL], logL[1];
    sample_U<G, DPL>(g, D, sk, u);
    forward_group<G, DPL, 1, FAM>(sm, g, row, u, scratch, logL);
    long long ne = 1;
    for (;;) {
        bool done;
        if (a.uniform_sampler) done = (logL[0] > contour) || (logL[0] == contour) || (ne >= 100);
        else done = !(logL[0] <= contour);
        if (done) break;
        sk = split_child(key, 1);
        key = split_child(key, 0);
        sample_U<G, DPL>(g, D, sk, u);
        forward_group<G, DPL, 1, FAM>(sm, g, row, u, scratch, logL);
        ne += 1;
    }
    const long long o = i - a.begin;
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D) a.out_U[o * D + j] = u[0][s];
    }
    if (g.lane == 0) {
        a.out_logL[o] = logL[0];
        a.out_nevals[o] = ne;
    }
}

template <int G, int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_draw(DrawArgs a) {
    extern __shared__ double smem[];
    NSB_FAMILY_SWITCH(a.model.family, draw_body<G, DPL, kFam>(a, smem));
}

// ---- vmap(Model.forward) / vmap(Model.transform) -----------------------------------------------
struct ForwardArgs {
    NsModelDesc model;
    const double *U;
    double *out_logL;
    double *out_X;
    long long n;
};

template <int G, int DPL, int FAM>
__device__ __forceinline__ void forward_body(const ForwardArgs &a, double *smem) {
    constexpr int DP = G * DPL;
    const int D = a.model.D;
    ModelSmem sm;
    stage_model<G, DPL>(a.model, smem, sm);
    __syncthreads();
    const Grp<G> g;
    constexpr int per_block = kThreadsPerBlock / G;
    const int local = threadIdx.x / G;
    const long long i = (long long) blockIdx.x * per_block + local;
    if (i >= a.n) return;
    DenseRow<G, DPL> row;
    row.load(sm, g.lane);
    double *scratch = smem + model_smem_doubles(sm.family, D, G, DPL, sm.K) + (size_t) local * DP;
    double u[1][DPL], X[1][DPL], logL[1];
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        u[0][s] = (j < D) ? a.U[i * D + j] : 0.5;
    }
    transform_dims<G, DPL, 1>(sm, g, u, X);
    loglik_group<G, DPL, 1, FAM>(sm, g, row, X, scratch, logL);
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int j = s * G + g.lane;
        if (j < D && a.out_X) a.out_X[i * D + j] = X[0][s];
    }
    if (g.lane == 0 && a.out_logL) a.out_logL[i] = logL[0];
}

template <int G, int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_forward(ForwardArgs a) {
    extern __shared__ double smem[];
    NSB_FAMILY_SWITCH(a.model.family, forward_body<G, DPL, kFam>(a, smem));
}

}  // namespace nsb

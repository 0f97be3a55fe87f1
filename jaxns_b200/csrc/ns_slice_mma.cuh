// Round-synchronous slice chains for the dense-Gaussian family with D <= 32: the quadratic forms of EIGHT
// proposals per warp run on the FP64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).
//
// Same algorithm, key tree and per-element arithmetic as k_slice_chains (ns_slice.cuh; reference:
// UniDimSliceSampler.get_sample_from_seed / _new_proposal, /root/reference/src/jaxns/samplers/uni_slice_sampler.py:114-273,
// :360-441, bases.py:63-75) -- only the mapping onto the machine differs:
//
//   k_slice_chains      one warp = one chain, lane = dimension.  Every evaluation exchanges the 32 residuals
//                       through shared memory, runs a 32-deep FMA row per lane (half of it multiplies the zeros
//                       above the diagonal) and ends in a 5-step shuffle butterfly: ~95 dependent FP64 warp
//                       instructions per evaluation, 166 registers, 2.7 warps per SM sub-partition at config 2
//                       => 42 % FP64 pipe, latency bound (profiles/r1/slice_r1_final.txt).
//   k_slice_chains_mma  one warp = 8 / P chains x P speculative proposals = the 8 columns of the B operand.
//                       With g = lane / 4 and t = lane % 4, lane (g, t) owns dimensions {4 s + t} of column g --
//                       exactly the B fragment layout of m8n8k4 -- so the prior transform of a round is 8
//                       INDEPENDENT quantiles per lane (ILP 8 instead of 1) and the residuals go into the MMA
//                       without any exchange.  L^-1 is kept as 8x4 A fragments (only the 20 tiles on or below
//                       the diagonal: 40 registers instead of 64), z = L^-1 r comes out as C fragments, and
//                       ||z||^2 needs a 3-step butterfly.  A DMMA occupies the FP64 pipe for 16 cycles = the
//                       rate of 8 DFMAs (profiles/r2/microbench_dmma.txt: 37.2 TFLOP/s, same pipe as DFMA), so
//                       the gain is not flops: it is the 32 LDS + 12 of 32 FMA rows + 2 butterfly steps per
//                       evaluation that disappear, and a warp that issues back-to-back independent work.
//
// All chains of a warp advance in lock-step ROUNDS: one round = every chain proposes its next P points (assuming
// the earlier ones of the round are rejected, ns_slice.cuh note (2)), all 8 are evaluated together, then each
// chain either shrinks its bracket or accepts and starts its next slice (bracket = line /\ unit cube) under a
// divergent branch.  Chains that have finished their S slices idle until the warp's last chain is done (~5 %).
//
// Streams (directions, proposal uniforms, continuation keys) come from k_chain_streams, as for k_slice_chains.
#pragma once
#include "ns_slice.cuh"

namespace nsb {

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// NB = row blocks of 8 (D <= 8 NB), P = speculative proposals per chain and round.
template <int NB, int P>
__global__ void __launch_bounds__(128) k_slice_chains_mma(SliceArgs a) {
    constexpr int KT = 2 * NB;                 // k tiles = register slots of a D-vector: dims j = 4 s + t
    constexpr int G = 4 * P;                   // lanes per chain
    constexpr int CH = 8 / P;                  // chains per warp
    constexpr int UPL = (kPre + G - 1) / G;    // proposal uniforms of the current slice held per lane
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);

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
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;     // B / C fragment coordinates: column g, k (dimension) residue t
    const int c = g / P, pme = g % P;          // chain slot inside the warp, proposal this 4-lane team evaluates
    const int li = lane & (G - 1);             // lane inside the chain
    const int cbase = c * G;                   // first lane of the chain
    const unsigned cmask = (G == 32) ? kFull : (((1u << (G & 31)) - 1u) << cbase);

    const int D = a.model.D, S = a.S, kph = a.k;
    const bool midpoint = a.midpoint != 0;
    const bool normal_prior = a.model.prior_kind == NSB200_PRIOR_NORMAL;
    const double *prm = a.model.params;        // [c, mu[D], Linv[D][D] row-major]
    const double lconst = __ldg(prm);
    const double contour = *contour_ptr;

    // ---- per-CTA constants in shared memory (prior, mean: read once per round, 3 x KT broadcast-free LDS.64) and
    // per-lane A fragments of L^-1 in registers (tiles on or below the diagonal).  Keeping the 3 KT prior / mean values
    // out of the register file is what leaves the scheduler room to interleave the KT quantile chains.
    __shared__ double s_mu[32], s_pa[32], s_pb[32];
    if (threadIdx.x < 32) {
        const int jj = threadIdx.x;
        const bool ok = jj < D;
        s_mu[jj] = ok ? __ldg(prm + 1 + jj) : 0.0;
        s_pa[jj] = ok ? __ldg(a.model.prior_a + jj) : 0.0;
        s_pb[jj] = ok ? __ldg(a.model.prior_b + jj) : 0.0;
    }
    __syncthreads();
    double A[NB * (NB + 1)];
    {
        int ai = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int s = 0; s < 2 * b + 2; ++s) {
                const int row = 8 * b + g, col = 4 * s + t;
                A[ai++] = (row < D && col < D && col <= row) ? __ldg(prm + 1 + D + (size_t) row * D + col) : 0.0;
            }
        }
    }

    const int warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long chain = a.chain_begin + ((long long) blockIdx.x * wpb + warp) * CH + c;
    bool fin = chain >= a.chain_end;           // finished, or no chain in this slot
    const long long out_row = chain - a.chain_begin;

    // ---- chain state (scalars replicated over the chain's lanes, vectors over its P teams)
    double U0[KT], d[KT], dn[KT];
    double uc[UPL], un[UPL];
    uint2 rkn = make_uint2(0, 0);
    Key rkey = Key{0, 0};
    double logL0 = 0.0, left = -1.0, right = 1.0, alpha = 1.0, alpha_n = 1.0;
    long long nev = 0;
    int j = 0, ne = 0;
#pragma unroll
    for (int s = 0; s < KT; ++s) {
        U0[s] = 0.5;
        d[s] = 0.0;
        dn[s] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < UPL; ++q) uc[q] = un[q] = 0.5;

    // stream of slice jn -> the "next" registers
    auto load_stream = [&](int jn) {
        const long long sn = out_row * S + jn;
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            const int jj = 4 * s + t;
            dn[s] = (jj < D) ? __ldg(a.pre_dirs + sn * D + jj) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < UPL; ++q) {
            const int n = li + q * G;
            un[q] = (n < kPre) ? __ldg(a.pre_us + sn * kPre + n) : 0.5;
        }
        rkn = __ldg(a.pre_rkeys + sn);
        alpha_n = a.alpha_tab ? __ldg(a.alpha_tab + jn) : alpha_schedule(jn, S);
    };
    // slice j starts: its stream moves from the "next" registers in, the loads of slice j + 1 are issued (their
    // latency hides behind this slice's rounds) and the bracket = line /\ unit cube is computed (_slice_bounds :41-64).
    // The P teams of the chain share the division work: team p takes the slots s = p (mod P).
    auto begin_slice = [&]() {
#pragma unroll
        for (int s = 0; s < KT; ++s) d[s] = dn[s];
#pragma unroll
        for (int q = 0; q < UPL; ++q) uc[q] = un[q];
        rkey = Key{rkn.x, rkn.y};
        alpha = alpha_n;
        if (j + 1 < S) load_stream(j + 1);
        // this team's slots s = pme, pme + P, ...; the NS reciprocals and quotients advance together (fast_div's
        // operations, interleaved over the slots)
        constexpr int NS = (KT + P - 1) / P;
        double us[NS], ds[NS], rc[NS], t1[NS], t0[NS];
        bool ok[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int s0 = i * P;
            us[i] = U0[s0];
            ds[i] = d[s0];
#pragma unroll
            for (int p = 1; p < P; ++p) {
                if (s0 + p < KT) {
                    us[i] = (pme == p) ? U0[s0 + p] : us[i];
                    ds[i] = (pme == p) ? d[s0 + p] : ds[i];
                }
            }
            ok[i] = (s0 + pme < KT) && (4 * (s0 + pme) + t < D);
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc[i]) : "d"(ds[i]));
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) rc[i] = fma(fma(-ds[i], rc[i], 1.0), rc[i], rc[i]);
#pragma unroll
        for (int i = 0; i < NS; ++i) rc[i] = fma(fma(-ds[i], rc[i], 1.0), rc[i], rc[i]);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double a1 = 1.0 - us[i], a0 = -us[i];
            double q1 = a1 * rc[i], q0 = a0 * rc[i];
            q1 = fma(fma(-ds[i], q1, a1), rc[i], q1);
            q0 = fma(fma(-ds[i], q0, a0), rc[i], q0);
            t1[i] = q1;
            t0[i] = q0;
        }
        // non-finite intermediates (d = 0, denormal d): fast_div falls back to the IEEE division -- a call, so that the
        // division sequence is not evaluated speculatively for every slot (never taken in practice)
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            if (ok[i] && !((t1[i] - t1[i] == 0.0) && (t0[i] - t0[i] == 0.0))) {
                if (!(t1[i] - t1[i] == 0.0)) t1[i] = slow_div(1.0 - us[i], ds[i]);
                if (!(t0[i] - t0[i] == 0.0)) t0[i] = slow_div(-us[i], ds[i]);
            }
        }
        // _slice_bounds :41-64: right = min over {t >= 0}, left = max over {t <= 0} of {t0_j, t1_j}.  For a point inside
        // the cube t0_j and t1_j have opposite signs, so per slot hi = max and lo = min are the two candidates and a
        // zero counts on both sides; the four-compare form is kept for the rounding-level case of a coordinate a hair
        // outside [0, 1] (uniform priors can accept one) -- selected per slot, so the common path is 2 min/max.
        double r = kInf, nl = kInf;  // right bound and MINUS the left bound, both >= 0
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double hi = fmax(t0[i], t1[i]), lo = fmin(t0[i], t1[i]);
            double ri = hi, li = -lo;
            if (!(hi > 0.0 && lo < 0.0)) {  // a zero, a NaN, or both on one side: the reference's four tests
                ri = kInf;
                li = kInf;
                if (t1[i] >= 0.0) ri = t1[i];
                if (t1[i] <= 0.0) li = -t1[i];
                if (t0[i] >= 0.0) ri = fmin(ri, t0[i]);
                if (t0[i] <= 0.0) li = fmin(li, -t0[i]);
            }
            if (ok[i]) {
                r = fmin(r, ri);
                nl = fmin(nl, li);
            }
        }
        right = group_min_nonneg(cmask, r + 0.0);
        left = -group_min_nonneg(cmask, nl + 0.0);
        ne = 0;
    };

#ifdef NSB_PROFILE
    unsigned long long prof[16];
    for (int k = 0; k < 16; ++k) prof[k] = 0;
#endif
    NSB_T0();
    // ---- chain prelude (bases.py:64; uni_slice_sampler.py:343-358, :410-413)
    if (!fin) {
        const Key chain_key = split_child(base_key, (uint64_t) chain);
        const Key seed_key = split_child(chain_key, 1);
        const double useed = uniform01(seed_key, 0);
        const long long sidx = seed_index(live_logL, a.seed_table, a.N, contour, useed);
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            const int jj = 4 * s + t;
            U0[s] = (jj < D) ? live_U[sidx * D + jj] : 0.5;
        }
        logL0 = live_logL[sidx];
        load_stream(0);
        begin_slice();
    }

    NSB_TICK(0)
    while (__any_sync(kFull, !fin)) {
#ifdef NSB_PROFILE
        prof[8] += 1;
#endif
        // ---- P proposals assuming each previous one is rejected (:92-111, :169-186)
        double ts[P], uus[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {  // the round's P uniforms first (independent shuffles), then the dependent chain
            const int n = ne + p;
            const int nn = n < kPre ? n : kPre - 1;
            if (UPL == 1) {
                uus[p] = __shfl_sync(kFull, uc[0], cbase + nn);
            } else {
                uus[p] = 0.0;
#pragma unroll
                for (int q = 0; q < UPL; ++q) {
                    const double v = __shfl_sync(kFull, uc[q], cbase + (nn & (G - 1)));
                    uus[p] = (nn / G == q) ? v : uus[p];
                }
            }
        }
        double l = left, r = right;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            double uu = uus[p];
            if (ne + p >= kPre) {  // beyond the precomputed uniforms: walk the run_key chain (:169), rare
                const Key t_key = split_child(rkey, 1);
                rkey = split_child(rkey, 0);
                uu = uniform01(t_key, 0);
            }
            const double tt = l + uu * (r - l);  // _pick_point_in_interval :83-85
            ts[p] = tt;
            if (tt < 0.0) l = midpoint ? alpha * tt : tt;  // _shrink_interval
            if (tt > 0.0) r = midpoint ? alpha * tt : tt;
        }
        double tm = ts[0];
#pragma unroll
        for (int p = 1; p < P; ++p) tm = (pme == p) ? ts[p] : tm;
        double x[KT], X[KT];
#pragma unroll
        for (int s = 0; s < KT; ++s) x[s] = fma(tm, d[s], U0[s]);
        NSB_TICK(1)
        // ---- prior transform of this lane's KT dimensions (wrapped_tfp_distribution.py:77-84)
        if (normal_prior) {
            double z[KT];
            ndtri_multi<KT>(x, z);
#pragma unroll
            for (int s = 0; s < KT; ++s) X[s] = z[s] * s_pb[4 * s + t] + s_pa[4 * s + t];
        } else {
#pragma unroll
            for (int s = 0; s < KT; ++s) X[s] = x[s] * s_pb[4 * s + t] + s_pa[4 * s + t];
        }
        NSB_TICK(2)
        // ---- z = L^-1 (X - mu) for the warp's 8 columns on the FP64 tensor path, then ||z||^2
        double sq0 = 0.0, sq1 = 0.0;
        {
            double rr[KT];
#pragma unroll
            for (int s = 0; s < KT; ++s) rr[s] = (4 * s + t < D) ? X[s] - s_mu[4 * s + t] : 0.0;
            int ai = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int s = 0; s < 2 * b + 2; ++s) dmma_8x8x4(c0, c1, A[ai++], rr[s]);
                sq0 = fma(c0, c0, sq0);
                sq1 = fma(c1, c1, sq1);
            }
        }
        // C fragment: rows g (+ 8 b), columns 2 t and 2 t + 1 -> sum over the rows = lanes with equal t
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            const double y0 = __shfl_xor_sync(kFull, sq0, o), y1 = __shfl_xor_sync(kFull, sq1, o);
            sq0 += y0;
            sq1 += y1;
        }
        double logL[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int col = c * P + p;  // column of proposal p of this lane's chain: held by the lanes with t = col / 2
            double qv;
            if (P % 2 == 0) {
                qv = __shfl_sync(kFull, (p & 1) ? sq1 : sq0, col >> 1);
            } else {
                const double v0 = __shfl_sync(kFull, sq0, col >> 1), v1 = __shfl_sync(kFull, sq1, col >> 1);
                qv = (col & 1) ? v1 : v0;
            }
            double ll = lconst - 0.5 * qv;
            if (ll != ll) ll = -kInf;  // ops.py:323-325
            logL[p] = ll;
        }
        NSB_TICK(3)
        // ---- first accepted proposal wins (:160-166)
        int hit = -1;
#pragma unroll
        for (int p = P - 1; p >= 0; --p) {
            const bool ok = (logL[p] > contour) || ((logL0 == contour) && (logL[p] == contour));
            if (ok) hit = p;
        }
#ifdef NSB_PROFILE
        prof[10] += __any_sync(kFull, !fin && hit >= 0) ? 1 : 0;
#endif
        if (!fin) {
            if (hit < 0 && ne + P > kMaxShrinkProposals) {
                // a bracket that has collapsed onto the seed point must accept (same point, same log L); it did
                // not -- non-deterministic or NaN likelihood.  Flag it and stay at the current point.
                if (a.err) atomicOr(a.err, NSB200_ERR_SHRINK_LOOP);
                hit = -2;
            }
            if (hit != -1) {
                if (hit >= 0) {
                    double th = ts[0], lh = logL[0];
#pragma unroll
                    for (int p = 1; p < P; ++p) {
                        th = (hit == p) ? ts[p] : th;
                        lh = (hit == p) ? logL[p] : lh;
                    }
#pragma unroll
                    for (int s = 0; s < KT; ++s) U0[s] = fma(th, d[s], U0[s]);
                    logL0 = lh;
                    nev += ne + hit + 1;
                } else {
                    nev += ne + P;
                }
                // phantom capture: cumulative_samples[-(k+1):-1] (:430-440)
                if (kph > 0 && j >= S - 1 - kph && j < S - 1 && pme == 0) {
                    const long long slot = out_row * kph + (j - (S - 1 - kph));
                    const long long pk = out_row * a.packed_row_doubles + (D + 2) + (long long) (j - (S - 1 - kph)) * (D + 1);
#pragma unroll
                    for (int s = 0; s < KT; ++s) {
                        const int jj = 4 * s + t;
                        if (jj < D) {
                            if (a.ph_U) a.ph_U[slot * D + jj] = U0[s];
                            packed_store(a, pk + jj, U0[s]);
                        }
                    }
                    if (t == 0) {
                        if (a.ph_logL) a.ph_logL[slot] = logL0;
                        packed_store(a, pk + D, logL0);
                    }
                }
                j += 1;
                if (j == S) {
                    fin = true;
                    left = -1.0;
                    right = 1.0;
                    if (pme == 0) {
#pragma unroll
                        for (int s = 0; s < KT; ++s) {
                            const int jj = 4 * s + t;
                            if (jj < D) {
                                if (a.out_U) a.out_U[out_row * D + jj] = U0[s];
                                packed_store(a, out_row * a.packed_row_doubles + jj, U0[s]);
                            }
                        }
                        if (t == 0) {
                            if (a.out_logL) a.out_logL[out_row] = logL0;
                            if (a.out_nevals) a.out_nevals[out_row] = nev;
                            packed_store(a, out_row * a.packed_row_doubles + D, logL0);
                            packed_store(a, out_row * a.packed_row_doubles + D + 1, __longlong_as_double(nev));
                        }
                    }
                    // a finished chain idles on the cube centre (central quantile region, no tail trips)
#pragma unroll
                    for (int s = 0; s < KT; ++s) {
                        U0[s] = 0.5;
                        d[s] = 0.0;
                    }
                } else {
                    begin_slice();
                }
            } else {
                ne += P;
                left = l;
                right = r;
            }
        }
        __syncwarp();
        NSB_TICK(4)
    }
#ifdef NSB_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < 16; ++k) atomicAdd(&g_prof[k], prof[k]);
#endif
}

}  // namespace nsb

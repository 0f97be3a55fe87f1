// Split slice step for likelihoods the library cannot fuse: the chain state lives in global memory
// and one slice move is cut into   propose -> [caller evaluates the batched likelihood on the device]
// -> accept/shrink/propose ...   (BASELINE north_star: "arbitrary user likelihoods are handled by
// splitting the slice step into propose and accept kernels around an on-device batched likelihood
// call, with no CPU fallback").
//
// The arithmetic and the key tree are the fused kernel's (ns_slice.cuh): same device functions, same
// order, so a split batch whose likelihood values equal the fused family's gives the same chains.
//
// Reference: BaseAbstractMarkovSampler._get_sample (/root/reference/src/jaxns/samplers/bases.py:63-75),
// UniDimSliceSampler.get_seed_point / get_sample_from_seed / _new_proposal
// (samplers/uni_slice_sampler.py:343-441, :114-273), Model.forward's NaN -> -inf rule
// (framework/ops.py:323-325), _single_uniform_sample (nested_samplers/common/uniform_sample.py:12-60).
#pragma once
#include "ns_slice.cuh"

namespace nsb {

// u32 per chain: j, ne, done, watchdog, run_key[2], after_key[2], sample_key2[2], need_grad, first_slice
constexpr int kSplitWords = 12;

// Proposals per chain and round.  A rejected proposal shrinks the bracket to its own t, which is known before its
// likelihood is, so the next proposals of a slice can be drawn assuming rejection and evaluated in the SAME batched
// likelihood call (first accepted wins) -- the fused kernel's speculation (ns_slice.cuh), bit-identical to one
// proposal per round.  The caller's likelihood is launch-bound at these batch sizes, so P x more rows per call cost
// nothing and the number of calls per slice drops from ~3.8 to ~1.1 at P = 8.
constexpr int kSplitMaxP = 8;

// Struct-of-arrays view of the per-chain state inside the caller's workspace.
struct SplitState {
    double *U0;       // [n, D] current point of the chain
    double *d;        // [n, D] direction of the current slice
    double *sc;       // [n, 4] left, right (bracket at the start of the round), unused, logL0
    double *ts;       // [n, kSplitMaxP] step lengths t of the round's proposals
    uint32_t *st;     // [n, kSplitWords]
    long long *nev;   // [n] likelihood evaluations so far
    double *phU;      // [n * k, D] phantom points
    double *phL;      // [n * k]
};

__host__ __device__ inline size_t split_align(size_t b) { return (b + 255) & ~(size_t) 255; }

__host__ inline size_t split_workspace_bytes(int D, long long n, int k) {
    size_t b = 0;
    b += 2 * split_align((size_t) n * D * 8);
    b += split_align((size_t) n * 4 * 8);
    b += split_align((size_t) n * kSplitMaxP * 8);
    b += split_align((size_t) n * kSplitWords * 4);
    b += split_align((size_t) n * 8);
    b += split_align((size_t) n * k * D * 8);
    b += split_align((size_t) n * k * 8);
    return b + 256;
}

__host__ inline SplitState split_state_view(void *ws, int D, long long n, int k) {
    char *p = (char *) (((uintptr_t) ws + 255) & ~(uintptr_t) 255);
    SplitState s;
    s.U0 = (double *) p;
    p += split_align((size_t) n * D * 8);
    s.d = (double *) p;
    p += split_align((size_t) n * D * 8);
    s.sc = (double *) p;
    p += split_align((size_t) n * 4 * 8);
    s.ts = (double *) p;
    p += split_align((size_t) n * kSplitMaxP * 8);
    s.st = (uint32_t *) p;
    p += split_align((size_t) n * kSplitWords * 4);
    s.nev = (long long *) p;
    p += split_align((size_t) n * 8);
    s.phU = (double *) p;
    p += split_align((size_t) n * k * D * 8);
    s.phL = (double *) p;
    return s;
}

struct SplitArgs {
    NsModelDesc model;  // prior transform only; the likelihood is the caller's
    Key key;
    const double *contour;
    const double *live_U;
    const double *live_logL;
    const double *seed_table;
    long long N;
    long long chain_begin, chain_end;
    int S, k, midpoint;
    const DevCtl *ctl;  // engine mode: key / contour / live buffer from the device-resident control block
    LiveSet live0, live1;
    SplitState state;
    int P;                            // proposals per chain and round (1..kSplitMaxP); row p * n + i = proposal p of chain i
    const double *prop_logL;          // [P, n] likelihood of the proposals written by the previous call
    double *prop_U;                   // [P, n, D] proposals in U space (out)
    double *prop_X;                   // [P, n, D] proposals through the prior transform (out, optional)
    unsigned long long *active;       // optional device counter: += chains that still need evaluations
    // gradient variants (uni_slice_sampler.py:202-214 gradient_slice = bit 0, :255-269 gradient_guided = bit 1): a chain
    // that starts a slice waits (need_grad) until the caller has evaluated d log L / dU at its current point
    // (k_split_export_U0 -> caller's autodiff -> mode 2)
    int grad_flags;
    const double *grad;               // mode 2: [n, D] gradient of log L w.r.t. U at the chains' current points
};

// mode 0: chain prelude (seed choice, first direction) + first proposal of slice 0
// mode 1: accept or shrink on prop_logL, then the next proposal (of this or the next slice)
// mode 2: (gradient variants) start the slice of every chain that waits for its gradient: Householder reflection of
//         the last direction at the accepted point (gradient_guided), gradient direction with left = 0 (gradient_slice)
template <int G, int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_split_step(SplitArgs a, int mode) {
    extern __shared__ double smem[];
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
    stage_model<G, DPL>(a.model, smem, sm);
    __syncthreads();
    const Grp<G> g;
    constexpr int per_block = kThreadsPerBlock / G;
    const long long row = (long long) blockIdx.x * per_block + threadIdx.x / G;
    const long long chain = a.chain_begin + row;
    if (chain >= a.chain_end) return;
    const double contour = *contour_ptr;
    const int S = a.S, kph = a.k;
    const bool midpoint = a.midpoint != 0;
    const int P = a.P;
    const long long nrows = a.chain_end - a.chain_begin;  // rows of one proposal plane
    uint32_t *st = a.state.st + row * kSplitWords;
    double *sc = a.state.sc + row * 4;
    double *tsv = a.state.ts + row * kSplitMaxP;

    double U0[DPL], d[DPL];
    int j, ne;
    Key run_key, after_key, sk2;
    double left, right, logL0;
    long long nev;
    bool new_slice;
    bool climb = false;  // gradient_slice with a usable gradient: only the uphill half of the bracket (left = 0)
    if (mode == 2) {
        if (st[2] || !st[10]) return;
        const bool first = st[11] != 0;
        j = (int) st[0];
        ne = 0;
        run_key = Key{st[4], st[5]};
        after_key = Key{st[6], st[7]};
        sk2 = Key{st[8], st[9]};
        left = right = 0.0;
        logL0 = sc[3];
        nev = a.state.nev[row];
        double gv[DPL];
        double ss = 0.0;
#pragma unroll
        for (int s = 0; s < DPL; ++s) {
            const int jj = s * G + g.lane;
            U0[s] = (jj < D) ? a.state.U0[row * D + jj] : 0.5;
            d[s] = (jj < D) ? a.state.d[row * D + jj] : 0.0;
            gv[s] = (jj < D) ? a.grad[row * D + jj] : 0.0;
            ss = fma(gv[s], gv[s], ss);
        }
        const double gn = sqrt(group_sum(g, ss));
        const bool finite = (gn - gn == 0.0);
        if (!first) {  // direction after the slice that just ended (:255-272)
            if (a.grad_flags & 2) {
                const Key after_key1 = split_child(after_key, 0);
                double rnd[DPL];
                sample_direction<G, DPL>(g, D, after_key1, rnd);
                const bool mask = !(gn >= 1e-10) || !finite;
                double dot = 0.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) dot = fma(d[s], gv[s] / gn, dot);
                dot = group_sum(g, dot);
                double refl[DPL], rs = 0.0;
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    refl[s] = d[s] - 2.0 * dot * (gv[s] / gn);
                    rs = fma(refl[s], refl[s], rs);
                }
                const double rn = sqrt(group_sum(g, rs));
#pragma unroll
                for (int s = 0; s < DPL; ++s) d[s] = mask ? rnd[s] : refl[s] / rn;
                nev += 1;
            } else {
                sample_direction<G, DPL>(g, D, after_key, d);
            }
        }
        if (a.grad_flags & 1) {  // climb the gradient (:202-214)
            nev += 1;
            const bool mask = (gn == 0.0) || !finite;
            if (!mask) {
#pragma unroll
                for (int s = 0; s < DPL; ++s) d[s] = gv[s] / gn;
                climb = true;
            }
        }
        new_slice = true;
    } else if (mode == 0) {
        // ---- chain prelude (bases.py:64; uni_slice_sampler.py:343-358, :410-413)
        const Key chain_key = split_child(base_key, (uint64_t) chain);
        const Key sample_key = split_child(chain_key, 0);
        const Key seed_key = split_child(chain_key, 1);
        const double useed = uniform01(seed_key, 0);
        const long long sidx = seed_index(live_logL, a.seed_table, a.N, contour, useed);
#pragma unroll
        for (int s = 0; s < DPL; ++s) {
            const int jj = s * G + g.lane;
            U0[s] = (jj < D) ? live_U[sidx * D + jj] : 0.5;
        }
        logL0 = live_logL[sidx];
        const Key direction_key = split_child(sample_key, 0);
        sk2 = split_child(sample_key, 1);
        sample_direction<G, DPL>(g, D, direction_key, d);
        j = 0;
        ne = 0;
        nev = 0;
        run_key = after_key = Key{0, 0};
        left = right = 0.0;
        new_slice = true;
        if (a.grad_flags) {  // the first slice starts in mode 2, once the gradient at the seed point is known
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int jj = s * G + g.lane;
                if (jj < D) {
                    a.state.U0[row * D + jj] = U0[s];
                    a.state.d[row * D + jj] = d[s];
                    a.prop_U[row * D + jj] = U0[s];
                }
            }
            if (g.lane == 0) {
                st[0] = 0u;
                st[1] = 0u;
                st[2] = 0u;
                st[3] = 0u;
                st[4] = st[5] = st[6] = st[7] = 0u;
                st[8] = sk2.a;
                st[9] = sk2.b;
                st[10] = 1u;
                st[11] = 1u;
                sc[0] = sc[1] = sc[2] = 0.0;
                sc[3] = logL0;
                a.state.nev[row] = 0;
                if (a.active) atomicAdd(a.active, 1ull);
            }
            return;
        }
    } else {
        if (st[10]) return;  // waits for its gradient (mode 2)
        if (st[2]) {  // chain finished: prop_U keeps its final point
            // a chain stopped by the shrink-loop watchdog keeps reporting it through bit 62 of the active counter
            if (st[3] && a.active && g.lane == 0) atomicOr(a.active, 1ull << 62);
            return;
        }
        j = (int) st[0];
        ne = (int) st[1];
        run_key = Key{st[4], st[5]};
        after_key = Key{st[6], st[7]};
        sk2 = Key{st[8], st[9]};
        left = sc[0];
        right = sc[1];
        logL0 = sc[3];
        nev = a.state.nev[row];
#pragma unroll
        for (int s = 0; s < DPL; ++s) {
            const int jj = s * G + g.lane;
            U0[s] = (jj < D) ? a.state.U0[row * D + jj] : 0.5;
            d[s] = (jj < D) ? a.state.d[row * D + jj] : 0.0;
        }
        // first accepted proposal of the round wins (:160-166)
        int hit = -1;
        double logL = 0.0;
        for (int p = P - 1; p >= 0; --p) {
            double v = a.prop_logL[(long long) p * nrows + row];
            if (v != v) v = -__longlong_as_double(0x7FF0000000000000ll);  // ops.py:323-325
            if ((v > contour) || ((logL0 == contour) && (v == contour))) {
                hit = p;
                logL = v;
            }
        }
        const bool ok = hit >= 0;
        if (ok) {
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int jj = s * G + g.lane;
                U0[s] = (jj < D) ? a.prop_U[((long long) hit * nrows + row) * D + jj] : 0.5;  // the point that was evaluated
            }
            logL0 = logL;
            nev += ne + hit + 1;
            // phantom capture: cumulative_samples[-(k+1):-1] (:430-440)
            if (kph > 0 && j >= S - 1 - kph && j < S - 1) {
                const long long slot = row * kph + (j - (S - 1 - kph));
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) a.state.phU[slot * D + jj] = U0[s];
                }
                if (g.lane == 0) a.state.phL[slot] = logL0;
            }
            j += 1;
            if (j == S) {
                if (a.grad_flags & 2) nev += 1;  // the reflection at the last accepted point is still evaluated (:257-258)
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) a.state.U0[row * D + jj] = U0[s];
                }
                if (g.lane == 0) {
                    st[0] = (uint32_t) j;
                    st[2] = 1u;
                    sc[3] = logL0;
                    a.state.nev[row] = nev;
                }
                return;
            }
            if (a.grad_flags) {  // the next slice starts in mode 2 with the gradient at the accepted point
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) {
                        a.state.U0[row * D + jj] = U0[s];
                        a.prop_U[row * D + jj] = U0[s];
                    }
                }
                if (g.lane == 0) {
                    st[0] = (uint32_t) j;
                    st[1] = 0u;
                    st[10] = 1u;
                    st[11] = 0u;
                    sc[3] = logL0;
                    a.state.nev[row] = nev;
                    if (a.active) atomicAdd(a.active, 1ull);
                }
                return;
            }
            sample_direction<G, DPL>(g, D, after_key, d);  // :272
            new_slice = true;
        } else {
            // every proposal of the round was rejected: _shrink_interval (:92-111) P times, then the next draws of
            // the while loop (:169-186)
            const double alpha = alpha_schedule(j, S);
            for (int p = 0; p < P; ++p) {
                const double t = tsv[p];
                if (t < 0.0) left = midpoint ? alpha * t : t;
                if (t > 0.0) right = midpoint ? alpha * t : t;
            }
            ne += P;
            new_slice = false;
            if (ne > kMaxShrinkProposals) {
                // watchdog (ns_slice.cuh kMaxShrinkProposals): the caller's likelihood is non-deterministic or NaN at
                // the chain's own seed point -- stop the chain where it is and raise NSB200_ERR_SHRINK_LOOP
                if (g.lane == 0) {
                    st[2] = 1u;
                    st[3] = 1u;
                    sc[3] = logL0;
                    a.state.nev[row] = nev + ne;
                    if (a.active) atomicOr(a.active, 1ull << 62);
                }
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) a.prop_U[row * D + jj] = U0[s];
                }
                return;
            }
        }
    }
    Key first_t_key = Key{0, 0};
    if (new_slice) {
        const Key slice_key = split_child(sk2, (uint64_t) j);  // :420
        run_key = split_child(slice_key, 0);                   // :201
        first_t_key = split_child(slice_key, 2);
        after_key = split_child(slice_key, 3);
        slice_bounds<G, DPL>(g, D, U0, d, left, right);
        if (climb) left = 0.0;  // :214
        ne = 0;  // proposals of this slice evaluated before this round
    }
    // the round's P proposals, each drawn as if the previous ones had been rejected (_pick_point_in_interval :83-85)
    {
        const double alpha = alpha_schedule(j, S);
        double l = left, r = right;
        for (int p = 0; p < P; ++p) {
            Key t_key;
            if (new_slice && p == 0) {
                t_key = first_t_key;
            } else {
                t_key = split_child(run_key, 1);  // :169
                run_key = split_child(run_key, 0);
            }
            const double uu = uniform01(t_key, 0);
            const double t = l + uu * (r - l);
            if (g.lane == 0) tsv[p] = t;
            if (t < 0.0) l = midpoint ? alpha * t : t;
            if (t > 0.0) r = midpoint ? alpha * t : t;
            double x[1][DPL];
            const long long prow = (long long) p * nrows + row;
#pragma unroll
            for (int s = 0; s < DPL; ++s) {
                const int jj = s * G + g.lane;
                x[0][s] = fma(t, d[s], U0[s]);
                if (jj < D) a.prop_U[prow * D + jj] = x[0][s];
            }
            if (a.prop_X) {
                double X[1][DPL];
                transform_dims<G, DPL, 1>(sm, g, x, X);
#pragma unroll
                for (int s = 0; s < DPL; ++s) {
                    const int jj = s * G + g.lane;
                    if (jj < D) a.prop_X[prow * D + jj] = X[0][s];
                }
            }
        }
    }
    if (new_slice) {
#pragma unroll
        for (int s = 0; s < DPL; ++s) {
            const int jj = s * G + g.lane;
            if (jj < D) {
                a.state.U0[row * D + jj] = U0[s];
                a.state.d[row * D + jj] = d[s];
            }
        }
    }
    if (g.lane == 0) {
        st[0] = (uint32_t) j;
        st[1] = (uint32_t) ne;
        st[2] = 0u;
        st[3] = 0u;
        st[10] = 0u;
        st[11] = 0u;
        st[4] = run_key.a;
        st[5] = run_key.b;
        st[6] = after_key.a;
        st[7] = after_key.b;
        st[8] = sk2.a;
        st[9] = sk2.b;
        sc[0] = left;
        sc[1] = right;
        sc[3] = logL0;
        a.state.nev[row] = nev;
        if (a.active) atomicAdd(a.active, 1ull);
    }
}

// Current points of the chains (where the caller evaluates d log L / dU for the gradient variants).
__global__ void k_split_export_U0(const DevCtl *ctl, SplitState s, long long n, int D, double *out_U) {
    if (ctl && !ctl->active) return;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < n * D; e += stride) out_U[e] = s.U0[e];
}

// Chain results out of the workspace: plain arrays (B1) and/or packed gather rows (engine).
__global__ void k_split_finish(const DevCtl *ctl, SplitState s, long long n, int D, int k, double *out_U,
                               double *out_logL, long long *out_nevals, double *ph_U, double *ph_logL, double *packed,
                               long long row_doubles) {
    if (ctl && !ctl->active) return;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < n * D; e += stride) {
        const long long r = e / D;
        const int jj = (int) (e - r * D);
        const double v = s.U0[e];
        if (out_U) out_U[e] = v;
        if (packed) packed[r * row_doubles + jj] = v;
        if (jj == 0) {
            const double l = s.sc[r * 4 + 3];
            const long long ne = s.nev[r];
            if (out_logL) out_logL[r] = l;
            if (out_nevals) out_nevals[r] = ne;
            if (packed) {
                packed[r * row_doubles + D] = l;
                packed[r * row_doubles + D + 1] = __longlong_as_double(ne);
            }
        }
    }
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < n * k * D; e += stride) {
        const long long pr = e / D;  // chain * k + slot
        const int jj = (int) (e - pr * D);
        const long long r = pr / k, slot = pr - r * k;
        const double v = s.phU[e];
        if (ph_U) ph_U[e] = v;
        if (packed) packed[r * row_doubles + (D + 2) + slot * (D + 1) + jj] = v;
        if (jj == 0) {
            const double l = s.phL[pr];
            if (ph_logL) ph_logL[pr] = l;
            if (packed) packed[r * row_doubles + (D + 2) + slot * (D + 1) + D] = l;
        }
    }
}

// Round `round` of _single_uniform_sample's redraw loop for prior draws [begin, end) of
// split(sample_key, N): U = sample_U(child1(child0^round(k_i))).  Rows with need[i] == 0 are left alone.
struct InitProposeArgs {
    NsModelDesc model;
    Key key;
    long long begin, end;
    int round;
    const unsigned char *need;  // optional [end - begin]
    double *out_U;
    double *out_X;  // optional
};

template <int G, int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_init_propose(InitProposeArgs a) {
    extern __shared__ double smem[];
    const int D = a.model.D;
    ModelSmem sm;
    stage_model<G, DPL>(a.model, smem, sm);
    __syncthreads();
    const Grp<G> g;
    constexpr int per_block = kThreadsPerBlock / G;
    const long long o = (long long) blockIdx.x * per_block + threadIdx.x / G;
    const long long i = a.begin + o;
    if (i >= a.end) return;
    if (a.need && !a.need[o]) return;
    Key key = split_child(a.key, (uint64_t) i);
    for (int r = 0; r < a.round; ++r) key = split_child(key, 0);
    const Key sk = split_child(key, 1);
    double u[1][DPL], X[1][DPL];
    sample_U<G, DPL>(g, D, sk, u);
    if (a.out_X) transform_dims<G, DPL, 1>(sm, g, u, X);
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int jj = s * G + g.lane;
        if (jj < D) {
            a.out_U[o * D + jj] = u[0][s];
            if (a.out_X) a.out_X[o * D + jj] = X[0][s];
        }
    }
}

// vmap(Model.transform) alone (no likelihood): used by the external-likelihood model.
template <int G, int DPL>
__global__ void __launch_bounds__(kThreadsPerBlock) k_transform(NsModelDesc model, const double *U, long long n, double *out_X) {
    extern __shared__ double smem[];
    const int D = model.D;
    ModelSmem sm;
    stage_model<G, DPL>(model, smem, sm);
    __syncthreads();
    const Grp<G> g;
    constexpr int per_block = kThreadsPerBlock / G;
    const long long i = (long long) blockIdx.x * per_block + threadIdx.x / G;
    if (i >= n) return;
    double u[1][DPL], X[1][DPL];
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int jj = s * G + g.lane;
        u[0][s] = (jj < D) ? U[i * D + jj] : 0.5;
    }
    transform_dims<G, DPL, 1>(sm, g, u, X);
#pragma unroll
    for (int s = 0; s < DPL; ++s) {
        const int jj = s * G + g.lane;
        if (jj < D) out_X[i * D + jj] = X[0][s];
    }
}

}  // namespace nsb

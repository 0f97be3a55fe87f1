// Engine-side kernels of the static nested-sampling loop: per-iteration key chain and bookkeeping,
// dead-store appends, rank-merge of the replaced shell, register update and termination decision.
//
// Reference: _main_ns_thread / _collect_shell / _add_samples_to_state
// (/root/reference/src/jaxns/nested_samplers/sharded/sharded_static.py:40-85, :210-324, :427-574),
// determine_termination (nested_samplers/common/termination.py:13-147),
// linear_to_log_stats / effective_sample_size_kish (internals/stats.py:55-86),
// replace_index = clamped dynamic_update_slice (internals/maps.py:15-25).
#pragma once
#include "ns_stats.cuh"
#include "ns_types.cuh"

namespace nsb {

// Global scratch of the cluster-wide register update (k_iter_epilogue).
struct EpiScratch {
    double gpart[3 * 3 * 64];  // CTA aggregates of the three group-wide scans (up to 64 CTAs, ev_gpart_stride)
    NsEvidenceCalc mid, fin;
    int not_plateau;
};

__device__ __forceinline__ long long clampll(long long v, long long lo, long long hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

// body prologue: the three key splits of one loop body (sharded_static.py:491,248,510), the contour
// (:250-251) and the dead-store bookkeeping of _add_samples_to_state (:54,76-78).
__global__ void k_iter_prologue(DevCtl *ctl, const NsRegister *reg, const LiveSet live0, const LiveSet live1,
                                long long m, long long kph, long long capacity, int intended_sender,
                                EpiScratch *epi, int no_speculation) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    (void) reg;
    (void) epi;
    // The loop condition this body runs under is the register of the body BEFORE the previous one (the host makes
    // this kernel wait for that update; done_iter = 0: the condition was already false at loop entry).  The previous
    // body's update may or may not have finished by now -- it is ignored on purpose, so that the decision does not
    // depend on timing: every rank of a multi-GPU run must take it identically (their chains meet at a barrier).
    {
        const long long dm = *(volatile long long *) &ctl->done_iter;
        const long long it = ctl->iteration + 1;
        // no_speculation: the host made this kernel wait for the previous body's update as well (runs whose store wraps
        // around cannot be rolled back: a speculative body would overwrite live rows of the ring)
        ctl->active = (dm >= 0 && (dm == 0 || dm <= it - 2 || no_speculation)) ? 0 : 1;
    }
    if (!ctl->active) return;
    {
        EpiJob &job = ctl->job[(ctl->iteration + 1) & 1];  // this body's slot (its previous user has been waited for)
        job.sum_new = 0;
        job.sum_live = 0;
        job.bar = 0;
        job.armed = 0;
    }
    const LiveSet &live = ctl->cur ? live1 : live0;
    Key k = split_child(ctl->key, 0);      // :491  key, ephemeral_key = split(state.key)
    ctl->sample_key = split_child(k, 1);   // :248  key, sample_key = split(state.key)
    k = split_child(k, 0);
    ctl->key = split_child(k, 0);          // :510  key, ephemeral_key = split(state.key)
    // the key chain does not depend on the data: the sample_key of the body after next is known already,
    // so its chain streams can be generated two bodies ahead, in the gaps this body leaves on the GPU
    {
        const Key k2 = split_child(split_child(split_child(ctl->key, 0), 0), 0);  // state key of body + 2
        ctl->stream_key[(ctl->body + 2) % 3] = split_child(split_child(k2, 0), 1);
        ctl->body += 1;
    }
    ctl->contour = live.logL[m - 1];
    ctl->disc_start = clampll(ctl->next_idx, 0, capacity - m);
    ctl->next_idx = (ctl->next_idx + m) % capacity;
    ctl->num_samples += m;
    ctl->sender = intended_sender ? ctl->next_idx : ctl->next_idx - 1;  // :261
    if (kph > 0) {
        ctl->ph_start = clampll(ctl->next_idx, 0, capacity - m * kph);
        ctl->next_idx = (ctl->next_idx + m * kph) % capacity;
        ctl->num_samples += m * kph;
    }
    ctl->iteration += 1;
}

// Append live rows [0, count) to the dead store at ctl->disc_start (discarded shell) or at an
// explicit offset (final live-set append, sharded_static.py:834-838).
__global__ void k_append_live(const DevCtl *ctl, const LiveSet live0, const LiveSet live1, DeadStore dead,
                              long long count, int D, int final_append) {
    if (!final_append && !ctl->active) return;
    const LiveSet &live = ctl->cur ? live1 : live0;
    const long long start = final_append ? clampll(ctl->next_idx, 0, dead.capacity - count) : ctl->disc_start;
    const long long total = count * D;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long) gridDim.x * blockDim.x) {
        dead.U[start * D + e] = live.U[e];
        if (e < count) {
            dead.sender[start + e] = live.sender[e];
            dead.logL[start + e] = live.logL[e];
            dead.nevals[start + e] = live.nevals[e];
            dead.phantom[start + e] = 0;
        }
    }
}

__global__ void k_finalize_ctl(DevCtl *ctl, long long count, long long capacity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->next_idx = (ctl->next_idx + count) % capacity;
    ctl->num_samples += count;
}

// Rank of every element of the next live set under a stable ascending sort of
// [new_0 .. new_{m-1}, survivor_m .. survivor_{N-1}]  (sharded_static.py:269-275): new rows come
// first on ties.  packed rows: [U[D], logL, nevals, ...].
// LANES = lanes cooperating on one element's count: 32 for small live sets (the kernel sits on the critical path
// between two slice kernels, more CTAs shorten it), 8 for large ones (every CTA re-reads all m new keys).
template <int kRankLanes>
__global__ void __launch_bounds__(256) k_merge_rank(const DevCtl *ctl, const LiveSet live0, const LiveSet live1,
                                                    const double *packed, long long row_doubles, int D,
                                                    long long m, long long N, unsigned *rank_out) {
    if (!ctl->active) return;
    __shared__ uint64_t tile[1024];
    const LiveSet &live = ctl->cur ? live1 : live0;
    const long long e = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / kRankLanes;
    const int sub = threadIdx.x % kRankLanes;
    const bool valid = e < N;
    const bool is_new = e < m;
    uint64_t key = 0;
    if (valid) key = sort_key_f64(is_new ? packed[e * row_doubles + D] : live.logL[e]);
    unsigned cnt = 0;
    for (long long t0 = 0; t0 < m; t0 += 1024) {
        const int tn = (int) min((long long) 1024, m - t0);
        __syncthreads();
        for (int q = threadIdx.x; q < tn; q += blockDim.x) tile[q] = sort_key_f64(packed[(t0 + q) * row_doubles + D]);
        __syncthreads();
        if (valid) {
            if (is_new) {
                for (int q = sub; q < tn; q += kRankLanes) {
                    const uint64_t kq = tile[q];
                    cnt += (kq < key) || (kq == key && (t0 + q) < e);
                }
            } else {
                for (int q = sub; q < tn; q += kRankLanes) cnt += (tile[q] <= key);
            }
        }
    }
#pragma unroll
    for (int o = kRankLanes >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (!valid || sub != 0) return;
    if (is_new) {
        // survivors strictly below (lower_bound over the sorted survivors)
        long long lo = m, hi = N;
        while (lo < hi) {
            long long mid = lo + ((hi - lo) >> 1);
            if (sort_key_f64(live.logL[mid]) < key) lo = mid + 1; else hi = mid;
        }
        cnt += (unsigned) (lo - m);
    } else {
        cnt += (unsigned) (e - m);
    }
    rank_out[e] = cnt;
}

// ---- rank merge for large shells: sorted tiles of the new keys + binary searches ---------------------------
// k_merge_rank compares every element with every new key: N * m compares (3.3e8 for the live set of an 8-GPU
// weak-scaling run: 175 us; 5e9 for config 5).  For m >= 2048 the new (key, index) pairs are sorted in tiles of
// 1024 (one CTA per tile, bitonic network in shared memory on the composite (key, index) order = the stable
// order); then, with T = ceil(m / 1024) tiles,
//   rank(new e)      = sum_t #{(k, i) in tile t : (k, i) < (key_e, e)} + #{survivors with key <  key_e}
//   rank(survivor p) = (p - m) + sum_t #{k in tile t : k <= key_p}
// are T ten-step binary searches per element, spread over 16 lanes: O(N T log 1024) instead of O(N m).
// (A single-CTA sort of all m keys was measured first: 280 us at m = 12800 -- one SM's issue rate.)
constexpr int kMergeTile = 1024;

__global__ void __launch_bounds__(kMergeTile) k_merge_sort_tiles(const DevCtl *ctl, const double *packed,
                                                                 long long row_doubles, int D, long long m,
                                                                 uint64_t *sorted_keys, unsigned *sorted_idx) {
    if (!ctl->active) return;
    __shared__ uint64_t sk[kMergeTile];
    __shared__ unsigned si[kMergeTile];
    const long long base = (long long) blockIdx.x * kMergeTile;
    {
        const long long r = base + threadIdx.x;
        sk[threadIdx.x] = (r < m) ? sort_key_f64(packed[r * row_doubles + D]) : 0xFFFFFFFFFFFFFFFFull;
        si[threadIdx.x] = (unsigned) r;  // pads carry indices >= m: they sort after NaN keys (same u64) of real rows
    }
    __syncthreads();
    for (int k = 2; k <= kMergeTile; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (threadIdx.x < (kMergeTile >> 1)) {
                const int t = threadIdx.x;
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // t with a zero bit inserted at log2(j)
                const int l = i | j;
                const uint64_t ka = sk[i], kb = sk[l];
                const unsigned ia = si[i], ib = si[l];
                const bool a_after_b = (ka > kb) || (ka == kb && ia > ib);
                const bool ascending = (i & k) == 0;
                if (a_after_b == ascending) {
                    sk[i] = kb;
                    sk[l] = ka;
                    si[i] = ib;
                    si[l] = ia;
                }
            }
            __syncthreads();
        }
    }
    sorted_keys[base + threadIdx.x] = sk[threadIdx.x];
    sorted_idx[base + threadIdx.x] = si[threadIdx.x];
}

constexpr int kMergeLanes = 16;  // lanes sharing the tiles of one element

__global__ void __launch_bounds__(256) k_merge_rank_tiles(const DevCtl *ctl, const LiveSet live0, const LiveSet live1,
                                                          const double *packed, long long row_doubles, int D,
                                                          long long m, long long N, const uint64_t *sorted_keys,
                                                          const unsigned *sorted_idx, unsigned *rank_out) {
    if (!ctl->active) return;
    const LiveSet &live = ctl->cur ? live1 : live0;
    const long long gt = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gt / kMergeLanes;
    const int sub = (int) (gt % kMergeLanes);
    const bool valid = e < N;
    const bool is_new = e < m;
    uint64_t key = 0;
    if (valid) key = sort_key_f64(is_new ? packed[e * row_doubles + D] : live.logL[e]);
    const int T = (int) ((m + kMergeTile - 1) / kMergeTile);
    unsigned cnt = 0;
    if (valid) {
        for (int t = sub; t < T; t += kMergeLanes) {
            const uint64_t *tk = sorted_keys + (long long) t * kMergeTile;
            const unsigned *ti = sorted_idx + (long long) t * kMergeTile;
            int lo = 0, hi = (int) min((long long) kMergeTile, m - (long long) t * kMergeTile);
            if (is_new) {  // entries strictly before (key, e) in the stable order
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const uint64_t km = tk[mid];
                    if (km < key || (km == key && ti[mid] < (unsigned) e)) lo = mid + 1; else hi = mid;
                }
            } else {  // new keys <= key (new rows come first on ties)
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (tk[mid] <= key) lo = mid + 1; else hi = mid;
                }
            }
            cnt += (unsigned) lo;
        }
    }
#pragma unroll
    for (int o = kMergeLanes >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (!valid || sub != 0) return;
    if (is_new) {
        long long lo = m, hi = N;  // survivors strictly below
        while (lo < hi) {
            const long long mid = lo + ((hi - lo) >> 1);
            if (sort_key_f64(live.logL[mid]) < key) lo = mid + 1; else hi = mid;
        }
        cnt += (unsigned) (lo - m);
    } else {
        cnt += (unsigned) (e - m);
    }
    rank_out[e] = cnt;
}

// Scatter rows into the other live buffer at their rank; phantom rows go to the dead store
// (add_phantom_samples_to_state, sharded_static.py:181-207).
__global__ void k_merge_scatter(DevCtl *ctl, const LiveSet live0, const LiveSet live1, const double *packed,
                                long long row_doubles, int D, long long m, long long N, int kph,
                                const unsigned *rank, DeadStore dead, int count_evals) {
    if (!ctl->active) return;
    const LiveSet &src = ctl->cur ? live1 : live0;
    const LiveSet &dst = ctl->cur ? live0 : live1;
    const long long total = N * D;
    long long sum_new = 0, sum_live = 0;  // n_evals of the new rows / of the merged live set (:301-304)
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long) gridDim.x * blockDim.x) {
        const long long e = t / D;
        const int j = (int) (t - e * D);
        const long long r = rank[e];
        const bool is_new = e < m;
        dst.U[r * D + j] = is_new ? packed[e * row_doubles + j] : src.U[e * D + j];
        if (j == 0) {
            long long nev;
            if (is_new) {
                nev = __double_as_longlong(packed[e * row_doubles + D + 1]);
                dst.sender[r] = ctl->sender;
                dst.logL[r] = packed[e * row_doubles + D];
                dst.logL_constraint[r] = ctl->contour;
                dst.nevals[r] = nev;
                sum_new += nev;
            } else {
                nev = src.nevals[e];
                dst.sender[r] = src.sender[e];
                dst.logL[r] = src.logL[e];
                dst.logL_constraint[r] = src.logL_constraint[e];
                dst.nevals[r] = nev;
            }
            sum_live += nev;
        }
    }
    if (count_evals) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum_new += __shfl_xor_sync(0xFFFFFFFFu, sum_new, o);
            sum_live += __shfl_xor_sync(0xFFFFFFFFu, sum_live, o);
        }
        if ((threadIdx.x & 31) == 0 && sum_live != 0) {
            EpiJob &job = ctl->job[ctl->iteration & 1];
            if (sum_new) atomicAdd(&job.sum_new, (unsigned long long) sum_new);
            atomicAdd(&job.sum_live, (unsigned long long) sum_live);
        }
    }
    if (kph > 0) {
        const long long ptotal = m * kph * (long long) D;
        for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < ptotal;
             t += (long long) gridDim.x * blockDim.x) {
            const long long row = t / D;  // phantom row index: chain * k + slot
            const int j = (int) (t - row * D);
            const long long chain = row / kph, slot = row - chain * kph;
            const double *p = packed + chain * row_doubles + (D + 2) + slot * (D + 1);
            const long long o = ctl->ph_start + row;
            dead.U[o * D + j] = p[j];
            if (j == 0) {
                dead.sender[o] = ctl->sender;
                dead.logL[o] = p[D];
                dead.nevals[o] = 0;
                dead.phantom[o] = 1;
            }
        }
    }
}

// Arrival barrier of the fused all-gather.  The slice kernel of every rank has stored its rows into every rank's
// gather buffer (NVLink peer stores); lane r of this one-warp kernel publishes this rank's arrival (epoch = body
// counter kept on the device, monotone over the engine's life) in rank r's flag array and waits for rank r's arrival in its own.
// The system-scope fence before the flag store orders the chains' peer stores (previous kernel in the stream)
// before it; the fence after the wait orders the merge kernels' reads after it.  A rank that does not arrive
// within ~30 s sets *err instead of hanging the GPU.  `force`: the run-entry barrier of engine_init (no rank may
// store rows of a new run into a peer that is still finishing the previous one).
struct PeerFlags {
    unsigned long long *p[8];
};

__global__ void k_peer_barrier(DevCtl *ctl, unsigned long long *epoch_dev, volatile unsigned long long *mine,
                               PeerFlags peers, int world, int me, int *err, int force, volatile long long *progress) {
    if (!force && !ctl->active) return;
    // The epoch counts the barriers this engine has passed.  It lives on the device and only advances for ACTIVE
    // bodies (and run entries), which every rank executes identically -- the host may enqueue different numbers of
    // no-op bodies per rank after the loop has ended without desynchronising the ranks.
    unsigned long long epoch = 0;
    if (threadIdx.x == 0) {
        epoch = *epoch_dev + 1;
        *epoch_dev = epoch;
    }
    epoch = __shfl_sync(0xFFFFFFFFu, epoch, 0);
    const int r = threadIdx.x;
    if (r >= world) return;
    // L_min agreement (north star: "L_min is agreed each iteration ... over NVLink"): the live log L is replicated, so
    // every rank derives the same contour without a collective; the barrier carries each rank's contour
    // along with its arrival flag and every rank checks all of them against its own -- an all-gather + compare of
    // L_min fused into the exchange that has to happen anyway.  A mismatch (a rank whose replicated state diverged)
    // raises NSB200_ERR_CONTOUR_MISMATCH instead of silently merging different runs.
    const unsigned long long my_contour = (unsigned long long) __double_as_longlong(ctl->contour);
    if (!force) *((volatile unsigned long long *) (peers.p[r] + 8 + me)) = my_contour;
    __threadfence_system();
    *((volatile unsigned long long *) (peers.p[r] + me)) = epoch;
    __threadfence_system();
    long long spins = 0;
    while (mine[r] < epoch) {
        __nanosleep(200);
        if (++spins > 150000000ll) {  // ~30 s
            atomicExch(err, 1);
            atomicOr(&ctl->err, NSB200_ERR_PEER_TIMEOUT);
            if (progress) {  // host-mapped: nsb200_engine_run polls it instead of spinning on a run that cannot finish
                progress[2] = 1;
                __threadfence_system();
            }
            break;
        }
    }
    __threadfence_system();
    if (!force && mine[r] == epoch && mine[8 + r] != my_contour) atomicOr(&ctl->err, NSB200_ERR_CONTOUR_MISMATCH);
}

// linear_to_log_stats (stats.py:55-74)
__device__ __forceinline__ void linear_to_log_stats(double log_f_mean, double log_f2_mean, double &mu, double &var) {
    mu = 2.0 * log_f_mean - 0.5 * log_f2_mean;
    const double d = log_f2_mean - 2.0 * log_f_mean;
    var = (d != d) ? d : fmax(d, 2.220446049250313e-16);  // jnp.maximum propagates NaN (-inf - -inf before any evidence), fmax would not
}

// determine_termination (termination.py:13-147)
__device__ inline void determine_termination(const NsTermCond &tc, NsRegister &reg) {
    long long reason = 0;
    bool done = false;
    const NsEvidenceCalc &ec = reg.evidence_calc, &ecr = reg.evidence_calc_with_remaining;
#define NSB_BIT(cond, bit) \
    if (cond) {            \
        done = true;       \
        reason += (1ll << (bit)); \
    }
    if (tc.mask & (1u << 4)) NSB_BIT((double) reg.num_samples_used >= tc.max_samples, 0)
    if (tc.mask & (1u << 1)) {
        double mu, var;
        linear_to_log_stats(ecr.log_Z_mean, ecr.log_Z2_mean, mu, var);
        NSB_BIT(var <= tc.evidence_uncert * tc.evidence_uncert, 1)
    }
    if (tc.mask & (1u << 3)) {
        double m1, v1, m0, v0;
        linear_to_log_stats(ecr.log_Z_mean, ecr.log_Z2_mean, m1, v1);
        linear_to_log_stats(ec.log_Z_mean, ec.log_Z2_mean, m0, v0);
        NSB_BIT((m1 - m0) < tc.dlogZ, 2)
    }
    if (tc.mask & (1u << 0)) {
        const double ess = exp(2.0 * ecr.log_Z_mean - ecr.log_dZ2_mean);
        NSB_BIT(ess >= tc.ess, 3)
    }
    if (tc.mask & (1u << 5)) NSB_BIT((double) reg.num_likelihood_evaluations >= tc.max_num_likelihood_evaluations, 4)
    if (tc.mask & (1u << 6)) NSB_BIT(reg.log_L_contour >= tc.log_L_contour, 5)
    if (tc.mask & (1u << 7)) NSB_BIT(reg.efficiency < tc.efficiency_threshold, 6)
    NSB_BIT(reg.plateau != 0, 7)
    if (tc.mask & (1u << 8)) NSB_BIT(reg.relative_spread < tc.rtol, 8)
    if (tc.mask & (1u << 9)) NSB_BIT(reg.absolute_spread < tc.atol, 9)
    NSB_BIT(reg.no_seed_points != 0, 10)
    if (tc.mask & (1u << 10)) {
        const double log_XL = ec.log_X_mean + ec.log_L;
        NSB_BIT(log_XL < reg.peak_log_XL + log(tc.peak_XL_frac), 11)
    }
#undef NSB_BIT
    reg.done = done ? 1 : 0;
    reg.termination_reason = reason;
}

// Register update of _collect_shell (sharded_static.py:281-323) followed by the loop condition.
// One thread-block cluster (kEvCluster CTAs x kEvThreads threads on neighbouring SMs, hardware
// cluster barrier between the scan phases).  `old` = live set before the merge (its first m rows are
// the discarded shell), `cur` = merged live set.
template <class Sync>
__device__ __forceinline__ void iter_epilogue_body(Sync &grp, DevCtl *ctl, NsRegister *reg, const LiveSet &live0,
                                                   const LiveSet &live1, const double *packed, long long row_doubles,
                                                   int D, long long m, long long N, const NsTermCond &tc, int init_only,
                                                   const double *tabT, const double *tabT2, const double *tabt,
                                                   long long tab_n, EpiScratch *epi, volatile long long *progress) {
    __shared__ double sh[3][34];
    const long long gtid = (long long) grp.rank() * blockDim.x + threadIdx.x;
    const long long nthreads = (long long) blockDim.x * grp.nranks();
    if (init_only) {
        // _main_ns_thread entry (:471-473): no_seed_points of the initial live set, then cond.
        if (gtid == 0) {
            const LiveSet &live = ctl->cur ? live1 : live0;
            reg->no_seed_points = live.logL[m - 1] >= live.logL[N - 1];
            determine_termination(tc, *reg);
            if (reg->done) ctl->done_iter = 0;
            if (progress) {
                progress[1] = reg->done;
                __threadfence_system();
                progress[0] = 0;
                __threadfence_system();
            }
        }
        return;
    }
    // the job armed by this body's k_iter_advance (at most one is armed: the host orders the next advance after
    // this kernel); none = the body was a no-op
    const int slot = ctl->job[0].armed ? 0 : (ctl->job[1].armed ? 1 : -1);
    if (slot < 0) return;
    EpiJob &job = ctl->job[slot];
    if (*(volatile long long *) &ctl->done_iter >= 0) {
        // the loop had already ended when this body started (it ran speculatively and will be rolled back); every CTA
        // takes this exit (done_iter was written by an earlier kernel), so no barrier is involved
        if (gtid == 0) job.armed = 0;
        return;
    }
    const LiveSet &old = job.old_cur ? live1 : live0;
    const LiveSet &cur = job.old_cur ? live0 : live1;
    EvSeq q;
    q.la = old.logL;
    q.na = nullptr;
    q.len_a = m;
    q.n_const_a = (double) N;  // :288, n = N for the whole shell
    q.lb = cur.logL;
    q.len_b = N;
    q.n_start_b = (double) N;  // :296, n = N..1
    q.tabT = tabT;
    q.tabT2 = tabT2;
    q.tabt = tabt;
    q.tab_n = tab_n;
    EvOut out;
    out.mid = &epi->mid;
    out.mark = m;
    out.fin = &epi->fin;
    out.per_sample = nullptr;
    if (m + N <= 8 * nthreads) evidence_scan_block<8>(q, reg->evidence_calc, out, sh, epi->gpart, grp);
    else evidence_scan_block<0>(q, reg->evidence_calc, out, sh, epi->gpart, grp);
    __threadfence();
    grp.sync();
    if (gtid == 0) {
        NsRegister r = *reg;
        const NsEvidenceCalc s_mid = epi->mid;
        const NsEvidenceCalc s_fin = epi->fin;
        r.num_samples_used = job.num_samples;
        r.evidence_calc = s_mid;
        r.evidence_calc_with_remaining = s_fin;
        r.num_likelihood_evaluations += (long long) *(volatile unsigned long long *) &job.sum_new;  // :301-302
        r.log_L_contour = job.contour;
        r.efficiency = (double) N / (double) (long long) *(volatile unsigned long long *) &job.sum_live;  // :304
        const double lo = cur.logL[0], hi = cur.logL[N - 1];
        r.plateau = (lo == hi) ? 1 : 0;  // :306 all(log_L == log_L[0]) on a sorted live set
        r.absolute_spread = fabs(hi - lo);
        r.relative_spread = 2.0 * r.absolute_spread / fabs(lo + hi);
        r.no_seed_points = cur.logL[m - 1] >= hi;
        r.peak_log_XL = fmax(r.peak_log_XL, s_mid.log_X_mean + s_mid.log_L);
        r.iteration = job.iteration;
        r.error_flags = *(volatile int *) &ctl->err;
        determine_termination(tc, r);
        if (r.error_flags) r.done = 1;  // a flagged run stops at once (the host turns the flags into an error)
        *reg = r;
        job.armed = 0;
        __threadfence();
        if (r.done) *(volatile long long *) &ctl->done_iter = r.iteration;
        if (progress) {
            // host-mapped pinned words: the host keeps a few bodies in flight and polls these instead of
            // synchronising the stream (progress[0] = bodies completed, progress[1] = done flag)
            progress[1] = r.done;
            __threadfence_system();
            progress[0] = r.iteration;
            __threadfence_system();
        }
    }
}

// End of a body on the main stream: freezes what the register update needs (it runs on its own stream while the
// next body starts), flips the live buffers and records the state a rollback would return to.
__global__ void k_iter_advance(DevCtl *ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!ctl->active) return;
    const int slot = (int) (ctl->iteration & 1);
    EpiJob &job = ctl->job[slot];
    job.num_samples = ctl->num_samples;
    job.iteration = ctl->iteration;
    job.contour = ctl->contour;
    job.old_cur = ctl->cur;
    ctl->cur ^= 1;
    CtlSnap &sn = ctl->snap[slot];
    sn.key = ctl->key;
    sn.next_idx = ctl->next_idx;
    sn.num_samples = ctl->num_samples;
    sn.iteration = ctl->iteration;
    sn.cur = ctl->cur;
    __threadfence();
    job.armed = 1;
}

// Discards the speculative body, if one ran (DevCtl::done_iter): the control block returns to the state after the
// body whose register ended the loop; k_blank_rows then resets the dead-store rows the speculative body appended.
__global__ void k_rollback(DevCtl *ctl, long long m, long long kph) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->spec_ran = 0;
    if (ctl->done_iter < 0 || ctl->iteration <= ctl->done_iter) return;
    ctl->spec_ran = 1;
    ctl->spec_disc_start = ctl->disc_start;
    ctl->spec_ph_start = kph > 0 ? ctl->ph_start : -1;
    (void) m;
    const CtlSnap &sn = ctl->snap[ctl->done_iter & 1];
    ctl->key = sn.key;
    ctl->next_idx = sn.next_idx;
    ctl->num_samples = sn.num_samples;
    ctl->iteration = sn.iteration;
    ctl->cur = sn.cur;
}

// create_init_state's empty rows (common/initialisation.py:38-45): sender 0, log L +inf, U 0, n_evals 0, not phantom
__global__ void k_blank_rows(const DevCtl *ctl, DeadStore dead, long long m, long long kph, int D) {
    if (!ctl->spec_ran) return;
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    for (int part = 0; part < 2; ++part) {
        const long long start = part == 0 ? ctl->spec_disc_start : ctl->spec_ph_start;
        const long long count = part == 0 ? m : m * kph;
        if (start < 0 || count <= 0) continue;
        for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < count * D;
             e += (long long) gridDim.x * blockDim.x) {
            dead.U[start * D + e] = 0.0;
            if (e < count) {
                dead.sender[start + e] = 0;
                dead.logL[start + e] = kInf;
                dead.nevals[start + e] = 0;
                dead.phantom[start + e] = 0;
            }
        }
    }
}

// Cluster form (8 CTAs on one GPC, hardware barrier) and grid form (8 CTAs anywhere, software barrier) of the
// register update; the engine launches the grid form inside the loop (NSB200_EPI_CLUSTER=1 switches back).
__global__ void __cluster_dims__(kEvCluster, 1, 1) __launch_bounds__(kEvThreads)
k_iter_epilogue(DevCtl *ctl, NsRegister *reg, const LiveSet live0, const LiveSet live1, const double *packed,
                long long row_doubles, int D, long long m, long long N, NsTermCond tc, int init_only,
                const double *tabT, const double *tabT2, const double *tabt, long long tab_n, EpiScratch *epi,
                volatile long long *progress) {
    ClusterSync grp;
    iter_epilogue_body(grp, ctl, reg, live0, live1, packed, row_doubles, D, m, N, tc, init_only, tabT, tabT2, tabt, tab_n,
                       epi, progress);
}

__global__ void __launch_bounds__(kEvThreads)
k_iter_epilogue_grid(DevCtl *ctl, NsRegister *reg, const LiveSet live0, const LiveSet live1, const double *packed,
                     long long row_doubles, int D, long long m, long long N, NsTermCond tc, int init_only,
                     const double *tabT, const double *tabT2, const double *tabt, long long tab_n, EpiScratch *epi,
                     volatile long long *progress) {
    GridSync grp{&ctl->job[ctl->job[0].armed ? 0 : 1].bar, 0u};
    iter_epilogue_body(grp, ctl, reg, live0, live1, packed, row_doubles, D, m, N, tc, init_only, tabT, tabT2, tabt, tab_n,
                       epi, progress);
}

}  // namespace nsb

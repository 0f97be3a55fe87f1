// Statistics kernels over the live / dead point sets: stable LSD radix argsort, rank-merge of a
// replaced shell into the sorted survivors, tree-structure live-point counting, the log-space
// evidence recurrences as parallel scans, logsumexp.
//
// Reference: count_crossed_edges (/root/reference/src/jaxns/internals/tree_structure.py:33-108),
// _update_evidence_calc_op / compute_evidence_stats (internals/shrinkage_statistics.py:43-157),
// cumulative_op_static/dynamic (internals/cumulative_ops.py:50-130), LogSpace.sum
// (internals/log_semiring.py:187-190), jnp.argsort call sites sharded_static.py:174,274.
#pragma once
#include <cooperative_groups.h>

#include "ns_math.cuh"
#include "../../include/nsb200.h"

namespace nsb {

// =================================================================================================
// Block-level scan helpers (warp shuffles + one shared array of 32 partials).
// =================================================================================================
struct OpAdd {
    __device__ __forceinline__ static double id() { return 0.0; }
    __device__ __forceinline__ static double ap(double a, double b) { return a + b; }
};
struct OpLae {
    __device__ __forceinline__ static double id() { return -__longlong_as_double(0x7FF0000000000000ll); }
    __device__ __forceinline__ static double ap(double a, double b) { return logaddexp(a, b); }
};

// Exclusive scan of one value per thread across the CTA; returns the exclusive prefix and (in
// `total`) the CTA aggregate.  `sh` = 33 doubles of shared memory.  Contains __syncthreads().
template <class Op>
__device__ __forceinline__ double block_scan_excl(double v, double *sh, double &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc = Op::ap(y, inc);
    }
    double exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) exc = Op::id();
    __syncthreads();
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = (lane < nw) ? sh[lane] : Op::id();
        double winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double y = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc = Op::ap(y, winc);
        }
        double wexc = __shfl_up_sync(0xFFFFFFFFu, winc, 1);
        if (lane == 0) wexc = Op::id();
        sh[lane] = wexc;
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    total = sh[32];
    return Op::ap(sh[warp], exc);
}

// Same, K values per thread scanned together: the K combines of a step are independent, so their
// latencies (logaddexp ~ 500 cycles) overlap instead of adding up over K separate scans.
template <class Op, int K>
__device__ __forceinline__ void block_scan_excl_k(const double (&v)[K], double (*sh)[34], double (&exc_out)[K]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double inc[K], exc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) inc[k] = v[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double y = __shfl_up_sync(0xFFFFFFFFu, inc[k], o);
            if (lane >= o) inc[k] = Op::ap(y, inc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        exc[k] = __shfl_up_sync(0xFFFFFFFFu, inc[k], 1);
        if (lane == 0) exc[k] = Op::id();
    }
    __syncthreads();
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < K; ++k) sh[k][warp] = inc[k];
    }
    __syncthreads();
    if (warp == 0) {
        double winc[K], w[K];
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = winc[k] = (lane < nw) ? sh[k][lane] : Op::id();
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double y = __shfl_up_sync(0xFFFFFFFFu, winc[k], o);
                if (lane >= o) winc[k] = Op::ap(y, winc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double wexc = __shfl_up_sync(0xFFFFFFFFu, winc[k], 1);
            if (lane == 0) wexc = Op::id();
            sh[k][lane] = wexc;
            if (lane == 31) sh[k][32] = winc[k];  // CTA aggregate
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) exc_out[k] = Op::ap(sh[k][warp], exc[k]);
}

// Exclusive scan across all threads of a thread-block CLUSTER (1..16 CTAs on neighbouring SMs):
// How the CTAs of a multi-CTA scan meet.  ClusterSync: a thread-block cluster and its hardware barrier (the CTAs
// must be placed on free SMs of ONE GPC).  GridSync: a plain grid of a few CTAs and a counter in global memory
// (zeroed before the launch) -- the CTAs can land on any free SM, which matters for the per-iteration register
// update: it runs next to the chain-stream generator, and a cluster waits until the generator has drained a GPC.
// The grid form spins, so it needs its CTAs to become resident eventually (they do: nothing else waits on them).
struct ClusterSync {
    __device__ __forceinline__ unsigned rank() const { return cooperative_groups::this_cluster().block_rank(); }
    __device__ __forceinline__ unsigned nranks() const { return cooperative_groups::this_cluster().num_blocks(); }
    __device__ __forceinline__ void sync() { cooperative_groups::this_cluster().sync(); }
};
struct GridSync {
    unsigned *counter;
    unsigned target;
    __device__ __forceinline__ unsigned rank() const { return blockIdx.x; }
    __device__ __forceinline__ unsigned nranks() const { return gridDim.x; }
    __device__ __forceinline__ void sync() {
        __syncthreads();
        target += gridDim.x;
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            while (*(volatile unsigned *) counter < target) {
            }
            __threadfence();
        }
        __syncthreads();
    }
};

// CTA-level scan, CTA aggregates exchanged through `gpart` ([ranks][K] doubles of global memory)
// around one barrier of the CTA group, prefix of the lower-ranked CTAs folded in by warp 0.
template <class Op, int K, class Sync>
__device__ __forceinline__ void cluster_scan_excl_k(const double (&v)[K], double (*sh)[34], double *gpart,
                                                    double (&out)[K], Sync &grp) {
    const unsigned rank = grp.rank(), nranks = grp.nranks();
    double exc[K];
    block_scan_excl_k<Op, K>(v, sh, exc);
    if (nranks == 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = exc[k];
        return;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) gpart[rank * K + k] = sh[k][32];
    }
    grp.sync();
    const int lane = threadIdx.x & 31;
    if (nranks > 16) {
        // many CTAs (final pass over a long dead set): every CTA scans all aggregates with a block scan and keeps
        // the exclusive prefix of its own rank (needs blockDim.x >= nranks)
        double agg[K], pre[K];
#pragma unroll
        for (int k = 0; k < K; ++k)
            agg[k] = (threadIdx.x < nranks) ? ((volatile double *) gpart)[threadIdx.x * K + k] : Op::id();
        __syncthreads();  // exc[] was combined from sh[k][warp] above: everyone is done reading it
        block_scan_excl_k<Op, K>(agg, sh, pre);
        __syncthreads();
        if (threadIdx.x == rank) {
#pragma unroll
            for (int k = 0; k < K; ++k) sh[k][33] = pre[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = Op::ap(sh[k][33], exc[k]);
        return;
    }
    if (threadIdx.x < 32) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double inc = ((unsigned) lane < nranks) ? ((volatile double *) gpart)[lane * K + k] : Op::id();
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                double y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc = Op::ap(y, inc);
            }
            double e = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
            if (lane == 0) e = Op::id();
            if ((unsigned) lane == rank) sh[k][33] = e;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = Op::ap(sh[k][33], exc[k]);
}

// =================================================================================================
// Stable LSD radix argsort on order-preserving u64 keys (8-bit digits, up to 8 passes), HBM-shaped:
//   * tiles of 4096 keys; every warp ranks a CONTIGUOUS 512-key chunk with __match_any_sync against warp-private
//     counters (no block barrier inside the ranking), one block scan turns the 8 x 256 warp counts into positions,
//     the tile is staged in shared memory in sorted order and written out in coalesced runs per digit;
//   * the 256 x tiles histogram table is scanned by 256 CTAs (one per digit row); the last CTA to finish turns the
//     row totals into digit bases -- no single-CTA scan over the whole table;
//   * a pass whose digit is the same for all keys is skipped on the device (log L keys share their top bits), and
//     an input that is already sorted skips every pass: the dead-point store of a k = 0 run IS sorted (each shell
//     is the sorted bottom of the live set and lies above the previous one), so count_crossed_edges' argsort
//     (tree_structure.py:39) degenerates to the identity there.  Which buffer holds the data after each pass is a
//     device-side decision, so all kernels pick their buffers through SortCtl.
// Algorithmic traffic per executed pass: 8 n (histogram read) + 12 n (read) + 12 n (write) bytes.
// =================================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096
constexpr int kSortWarps = kSortThreads / 32;

struct SortCtl {
    int curs[9];        // buffer (0 / 1) holding the data at the start of pass p; curs[8] = final
    int skip[8];        // pass p is the identity (one digit value holds all keys)
    int sorted;         // the input was already in stable ascending order
    unsigned done[8];   // CTAs of the row scan that have finished (last one computes the digit bases)
    unsigned total[256];
    unsigned base[256];
};

__global__ void k_sort_init(SortCtl *ctl) {
    const int t = threadIdx.x;
    if (t < 9) ctl->curs[t] = 0;
    if (t < 8) {
        ctl->skip[t] = 0;
        ctl->done[t] = 0;
    }
    if (t == 0) ctl->sorted = 1;
}

// keys_out[i + offset] = sort_key(x[i]); vals = iota.  `lead_neg_inf` prepends the -inf root node.
__global__ void k_sort_prep(const double *x, long long n, int lead_neg_inf, uint64_t *keys, uint32_t *vals,
                            SortCtl *ctl) {
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = n + (lead_neg_inf ? 1 : 0);
    if (i >= total) return;
    const double kNegInf = -__longlong_as_double(0x7FF0000000000000ll);
    double v, vp = kNegInf;
    if (lead_neg_inf) {
        v = (i == 0) ? kNegInf : x[i - 1];
        if (i > 1) vp = x[i - 2];
    } else {
        v = x[i];
        if (i > 0) vp = x[i - 1];
    }
    const uint64_t k = sort_key_f64(v);
    keys[i] = k;
    vals[i] = (uint32_t) i;
    if (i > 0 && sort_key_f64(vp) > k) ctl->sorted = 0;  // benign race: every writer stores 0
}

__device__ __forceinline__ long long sort_item_index(long long tile_base, int warp, int r, int lane) {
    return tile_base + (long long) warp * (kSortItems * 32) + r * 32 + lane;
}

__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint64_t *keys0, const uint64_t *keys1,
                                                             const SortCtl *ctl, long long n, int pass,
                                                             uint32_t *hist, int nblocks) {
    if (ctl->sorted) return;
    __shared__ uint32_t h[256];
    const uint64_t *keys = ctl->curs[pass] ? keys1 : keys0;
    const int shift = pass * 8;
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long) blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const long long i = base + (long long) r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t) threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// CTA b: in-place exclusive scan of row b of the digit-major table (`cnt` entries); the last CTA to finish turns the
// 256 row totals into digit bases and decides whether the pass is the identity.
__global__ void __launch_bounds__(256) k_radix_scan_rows(uint32_t *hist, SortCtl *ctl, long long n, int pass, int cnt) {
    if (ctl->sorted) return;
    __shared__ uint32_t sh[9];
    __shared__ int last;
    uint32_t *row = hist + (size_t) blockIdx.x * cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int c0 = 0; c0 < cnt; c0 += 256) {
        const int i = c0 + threadIdx.x;
        const uint32_t v = (i < cnt) ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        __syncthreads();
        if (lane == 31) sh[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t x = sh[w];
            if (w < warp) woff += x;
            tot += x;
        }
        if (i < cnt) row[i] = carry + woff + inc - v;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        ctl->total[blockIdx.x] = carry;
        __threadfence();
        last = atomicAdd(&ctl->done[pass], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const uint32_t v = ((volatile unsigned *) ctl->total)[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (w < warp) woff += sh[w];
    ctl->base[threadIdx.x] = woff + inc - v;
    const int trivial = __syncthreads_or(v == (uint32_t) n);
    if (threadIdx.x == 0) {
        ctl->skip[pass] = trivial;
        ctl->curs[pass + 1] = ctl->curs[pass] ^ (trivial ? 0 : 1);
    }
}

__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(uint64_t *keys0, uint64_t *keys1, uint32_t *vals0,
                                                                uint32_t *vals1, const SortCtl *ctl, long long n,
                                                                int pass, const uint32_t *offsets, int nblocks) {
    if (ctl->sorted || ctl->skip[pass]) return;
    extern __shared__ unsigned char rs_smem[];
    uint64_t *skeys = (uint64_t *) rs_smem;                               // [kSortTile]
    uint32_t *svals = (uint32_t *) (skeys + kSortTile);                   // [kSortTile]
    uint32_t *wcnt = svals + kSortTile;                                   // [kSortWarps][256]
    uint32_t *tstart = wcnt + kSortWarps * 256;                           // [256] first sorted position of a digit
    uint32_t *gbase = tstart + 256;                                       // [256] global position of that run
    __shared__ uint32_t sh[9];
    const int cur = ctl->curs[pass];
    const uint64_t *keys_in = cur ? keys1 : keys0;
    const uint32_t *vals_in = cur ? vals1 : vals0;
    uint64_t *keys_out = cur ? keys0 : keys1;
    uint32_t *vals_out = cur ? vals0 : vals1;
    const int shift = pass * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) wcnt[w * 256 + tid] = 0;
    gbase[tid] = ctl->base[tid] + offsets[(size_t) tid * nblocks + blockIdx.x];
    __syncthreads();
    const long long tile = (long long) blockIdx.x * kSortTile;
    uint64_t key[kSortItems];
    uint32_t val[kSortItems];
    uint32_t meta[kSortItems];  // digit | rank inside the warp's chunk << 9
    uint32_t *mycnt = wcnt + warp * 256;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const long long i = sort_item_index(tile, warp, r, lane);
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0;
        val[r] = valid ? vals_in[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const long long i = sort_item_index(tile, warp, r, lane);
        const bool valid = i < n;
        const unsigned digit = valid ? (unsigned) ((key[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
        const unsigned rk = __popc(peers & ((1u << lane) - 1u));
        uint32_t before = 0;
        if (valid) before = mycnt[digit];
        __syncwarp();
        if (valid && rk == 0) mycnt[digit] = before + __popc(peers);
        __syncwarp();
        meta[r] = digit | ((before + rk) << 9);
    }
    __syncthreads();
    // digit tid: warp counts -> exclusive prefix over the warps, total -> exclusive scan over the digits
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = wcnt[w * 256 + tid];
            wcnt[w * 256 + tid] = run;
            run += c;
        }
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) sh[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
            if (w < warp) woff += sh[w];
        tstart[tid] = woff + inc - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const unsigned digit = meta[r] & 511u;
        if (digit < 256u) {
            const uint32_t pos = tstart[digit] + mycnt[digit] + (meta[r] >> 9);
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }
    __syncthreads();
    const int count = (int) min((long long) kSortTile, n - tile);
    for (int q = tid; q < count; q += kSortThreads) {
        const uint64_t k = skeys[q];
        const unsigned digit = (unsigned) ((k >> shift) & 255u);
        const uint32_t dst = gbase[digit] + ((uint32_t) q - tstart[digit]);
        keys_out[dst] = k;
        vals_out[dst] = svals[q];
    }
}

constexpr size_t kSortScatterSmem = (size_t) kSortTile * 12 + (size_t) kSortWarps * 256 * 4 + 2 * 256 * 4;

struct SortWorkspace {
    uint64_t *keys[2];
    uint32_t *vals[2];
    uint32_t *hist;
    SortCtl *ctl;
    int nblocks;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t) 255; }

inline size_t sort_workspace_bytes(long long n) {
    const long long nb = (n + kSortTile - 1) / kSortTile;
    return 2 * align256((size_t) n * 8) + 2 * align256((size_t) n * 4) + align256((size_t) 256 * nb * 4) +
           align256(sizeof(SortCtl)) + 256;
}

inline SortWorkspace carve_sort_workspace(void *ws, long long n) {
    SortWorkspace w;
    char *p = (char *) ws;
    p = (char *) align256((size_t) p);
    w.nblocks = (int) ((n + kSortTile - 1) / kSortTile);
    w.keys[0] = (uint64_t *) p; p += align256((size_t) n * 8);
    w.keys[1] = (uint64_t *) p; p += align256((size_t) n * 8);
    w.vals[0] = (uint32_t *) p; p += align256((size_t) n * 4);
    w.vals[1] = (uint32_t *) p; p += align256((size_t) n * 4);
    w.hist = (uint32_t *) p; p += align256((size_t) 256 * w.nblocks * 4);
    w.ctl = (SortCtl *) p;
    return w;
}

// Sorts keys[0]/vals[0] (n entries, written by k_sort_prep after k_sort_init); the result is in
// keys[ctl->curs[8]] / vals[ctl->curs[8]].  Returns a CUDA error code.
inline cudaError_t radix_sort_pairs(const SortWorkspace &w, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int) kSortScatterSmem);
        if (err != cudaSuccess) return err;
        attr_set = true;
    }
    for (int pass = 0; pass < 8; ++pass) {
        k_radix_hist<<<w.nblocks, kSortThreads, 0, st>>>(w.keys[0], w.keys[1], w.ctl, n, pass, w.hist, w.nblocks);
        k_radix_scan_rows<<<256, 256, 0, st>>>(w.hist, w.ctl, n, pass, w.nblocks);
        k_radix_scatter<<<w.nblocks, kSortThreads, kSortScatterSmem, st>>>(w.keys[0], w.keys[1], w.vals[0], w.vals[1],
                                                                            w.ctl, n, pass, w.hist, w.nblocks);
    }
    return cudaGetLastError();
}

__global__ void k_vals_to_i64(const uint32_t *vals0, const uint32_t *vals1, const SortCtl *ctl, long long n,
                              long long *out) {
    const uint32_t *vals = ctl->curs[8] ? vals1 : vals0;
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (long long) vals[i];
}

// =================================================================================================
// Tree-structure live-point counting.
// =================================================================================================
// out_degree[sender] += 1 (tree_structure.py:54-56).  All replacements of a shell share one sender, so the
// increments come in runs of thousands on one address: equal values inside a warp are combined first.
__global__ void k_out_degree(const long long *sender, long long M, int *outdeg) {
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < M;
    long long s = valid ? sender[i] : -1;
    if (valid && s < 0) s = 0;  // lax.max(sender, 0)
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, s);
    if (valid && __popc(peers & ((1u << lane) - 1u)) == 0) atomicAdd(&outdeg[s], __popc(peers));
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// pass 1: per-tile sums of delta_i = outdeg[sort_idx[i]] - 1
__global__ void __launch_bounds__(kScanThreads) k_tree_tile_sums(const uint32_t *idx0, const uint32_t *idx1,
                                                                 const SortCtl *ctl, const int *outdeg, long long n,
                                                                 int *tile_sums) {
    __shared__ int sh[kScanThreads / 32];
    const uint32_t *sort_idx = ctl->curs[8] ? idx1 : idx0;
    const long long base = (long long) blockIdx.x * kScanTile + (long long) threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        const long long i = base + r;
        if (i < n) s += outdeg[sort_idx[i]] - 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += sh[w];
        tile_sums[blockIdx.x] = t;
    }
}

// In-place exclusive scan of `cnt` uint32 entries by one CTA of 1024 threads (tile sums: n / 4096 entries).
__global__ void __launch_bounds__(1024) k_scan_u32_excl(uint32_t *data, long long cnt) {
    __shared__ uint32_t sh[33];
    const long long per = (cnt + blockDim.x - 1) / blockDim.x;
    const long long b = (long long) threadIdx.x * per, e = min(cnt, b + per);
    uint32_t s = 0;
    for (long long i = b; i < e; ++i) s += data[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = sh[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += y;
        }
        sh[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = sh[warp] + (inc - s);
    for (long long i = b; i < e; ++i) {
        uint32_t v = data[i];
        data[i] = run;
        run += v;
    }
}

// pass 3: per-tile inclusive scan + outputs
__global__ void __launch_bounds__(kScanThreads) k_tree_apply(const uint32_t *idx0, const uint32_t *idx1,
                                                             const SortCtl *ctl, const int *outdeg, long long n,
                                                             const int *tile_offsets, long long M,
                                                             long long num_samples, long long *out_idx,
                                                             int *out_nlive) {
    __shared__ int sh[kScanThreads / 32 + 1];
    const uint32_t *sort_idx = ctl->curs[8] ? idx1 : idx0;
    const long long base = (long long) blockIdx.x * kScanTile + (long long) threadIdx.x * kScanItems;
    int loc[kScanItems];
    int s = 0;
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        const long long i = base + r;
        int dlt = 0;
        if (i < n) dlt = outdeg[sort_idx[i]] - 1;
        s += dlt;
        loc[r] = s;
    }
    // exclusive scan of s across the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) {
            int v = sh[w];
            sh[w] = run;
            run += v;
        }
    }
    __syncthreads();
    const int prefix = tile_offsets[blockIdx.x] + sh[warp] + (inc - s) + 1;  // init = 1
    const bool dynamic = num_samples >= 0;
    const int fake = dynamic ? (int) (M - num_samples) : 0;
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        const long long i = base + r;
        if (i < M) {  // drop the last node (crossed[:-1]); samples_indices = sort_idx[1:] - 1
            int crossed = prefix + loc[r];
            if (dynamic) crossed = ((i < num_samples) ? crossed : fake) - fake;
            out_nlive[i] = crossed;
            out_idx[i] = (long long) sort_idx[i + 1] - 1;
        }
    }
}

// =================================================================================================
// Evidence recurrences as scans.  One CTA; every thread owns a contiguous chunk.
// =================================================================================================
// The scanned sequence is segment A followed by segment B:
//   A: log_L = la[i], n = na ? na[i] : n_const_a          (dead points / discarded shell)
//   B: log_L = lb[i], n = n_start_b - i                   (remaining live points, n = N..1)
struct EvSeq {
    const double *la;
    const double *na;
    long long len_a;
    double n_const_a;
    const double *lb;
    long long len_b;
    double n_start_b;
    // optional tables of the n-dependent terms for integer n in [1, tab_n]: T, T2 and t = -log(n+1)
    // (t is valid up to tab_n + 1); built once per engine by k_ev_tables
    const double *tabT;
    const double *tabT2;
    const double *tabt;
    long long tab_n;
};

struct EvTerms {
    double T, T2, t, t2, tT, mid;
};

__device__ __forceinline__ void ev_get(const EvSeq &q, long long i, double &logL, double &n) {
    if (i < q.len_a) {
        logL = q.la[i];
        n = q.na ? q.na[i] : q.n_const_a;
    } else {
        const long long j = i - q.len_a;
        logL = q.lb[j];
        n = q.n_start_b - (double) j;
    }
}

__device__ __forceinline__ EvTerms ev_terms(const EvSeq &q, double logL, double prevL, double n) {
    const double kLog2 = 0.6931471805599453, kLogHalf = -0.6931471805599453;
    EvTerms e;
    if (q.tabT) {
        const long long ni = (long long) n;
        if ((double) ni == n && ni >= 1 && ni <= q.tab_n) {
            const double tn = q.tabt[ni], tn1 = q.tabt[ni + 1];
            e.mid = kLogHalf + logaddexp(logL, prevL);
            e.T = q.tabT[ni];
            e.T2 = q.tabT2[ni];
            e.t = tn;
            e.t2 = kLog2 + tn + tn1;
            e.tT = e.T + tn1;
            return e;
        }
    }
    const double ln = log(n), lnp1 = log(n + 1.0), lnp2 = log(n + 2.0);
    e.mid = kLogHalf + logaddexp(logL, prevL);
    e.T = -logaddexp(0.0, -ln);
    e.t = -lnp1;
    e.T2 = -logaddexp(0.0, kLog2 - ln);
    e.t2 = kLog2 - lnp1 - lnp2;
    e.tT = e.T - lnp2;
    return e;
}

// T[n] = -logaddexp(0, -log n), T2[n] = -logaddexp(0, log 2 - log n), t[n] = -log(n + 1), n = 0 .. nmax+1
__global__ void k_ev_tables(long long nmax, double *tabT, double *tabT2, double *tabt) {
    const double kLog2 = 0.6931471805599453;
    for (long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x; n <= nmax + 1;
         n += (long long) gridDim.x * blockDim.x) {
        const double ln = log((double) n);
        tabT[n] = -logaddexp(0.0, -ln);
        tabT2[n] = -logaddexp(0.0, kLog2 - ln);
        tabt[n] = -log((double) n + 1.0);
    }
}

// doubles of global scratch per cluster-level scan (K <= 3 values per rank); 16 ranks -> the historic 48
__host__ __device__ inline size_t ev_gpart_stride(unsigned nranks) { return 3 * (size_t) (nranks < 16 ? 16 : nranks); }

struct EvOut {
    NsEvidenceCalc *mid;     // state after element index mark-1 (nullable)
    long long mark;
    NsEvidenceCalc *fin;     // state after the last element (nullable)
    double *per_sample;      // [8][M] field-major (nullable)
};

// Must be called by all threads of a CTA.  sh = 33 doubles.  CACHE > 0: every thread owns at most
// CACHE elements and keeps their terms (3 log + 3 logaddexp each) in registers across the passes --
// the per-iteration register update (m + N elements over 1024 threads) runs this way.
// `gpart` = 3 * 16 * 3 doubles of global scratch for the cluster-level scans (nullptr is fine for a
// single CTA).
template <int CACHE, class Sync>
__device__ inline void evidence_scan_block(const EvSeq &q, const NsEvidenceCalc &init, const EvOut &out, double (*sh)[34],
                                           double *gpart, Sync &grp) {
    const double kLog2 = 0.6931471805599453;
    const long long M = q.len_a + q.len_b;
    const long long nthreads = (long long) blockDim.x * grp.nranks();
    const long long gtid = (long long) grp.rank() * blockDim.x + threadIdx.x;
    const long long per = (M + nthreads - 1) / nthreads;
    const long long b = min(M, gtid * per), e = min(M, b + per);
    double prev0 = init.log_L;
    if (b > 0 && b < M) {
        double nn;
        ev_get(q, b - 1, prev0, nn);
    }
    constexpr int C = CACHE > 0 ? CACHE : 1;
    EvTerms ct[C];
    double cl[C];
    if (CACHE > 0) {
        double prevL = prev0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const long long i = b + c;
            if (i < e) {
                double logL, n;
                ev_get(q, i, logL, n);
                ct[c] = ev_terms(q, logL, prevL, n);
                cl[c] = logL;
                prevL = logL;
            }
        }
    }
#define NSB_EV_LOOP(BODY)                                             \
    if (CACHE > 0) {                                                  \
        _Pragma("unroll") for (int c = 0; c < C; ++c) {               \
            const long long i = b + c;                                \
            if (i < e) {                                              \
                const EvTerms t = ct[c];                              \
                const double logL = cl[c];                            \
                BODY                                                  \
            }                                                         \
        }                                                             \
    } else {                                                          \
        double prevL = prev0;                                         \
        for (long long i = b; i < e; ++i) {                           \
            double logL, n;                                           \
            ev_get(q, i, logL, n);                                    \
            const EvTerms t = ev_terms(q, logL, prevL, n);            \
            prevL = logL;                                             \
            BODY                                                      \
        }                                                             \
    }
    // pass 1: cumsum of T, T2
    double sT = 0.0, sT2 = 0.0;
    NSB_EV_LOOP({ (void) i; (void) logL; sT += t.T; sT2 += t.T2; })
    double in2[2] = {sT, sT2}, ex2[2];
    cluster_scan_excl_k<OpAdd, 2>(in2, sh, gpart, ex2, grp);
    const double X0 = init.log_X_mean + ex2[0];
    const double X20 = init.log_X2_mean + ex2[1];
    // pass 2: Z, dZ2, W = ZX / X
    const double kNegInf = OpLae::id();
    double sa = kNegInf, sb = kNegInf, sw = kNegInf;
    {
        double lX = X0, lX2 = X20;
        NSB_EV_LOOP({
            (void) i; (void) logL;
            sa = logaddexp(sa, lX + t.t + t.mid);
            sb = logaddexp(sb, lX2 + t.t2 + 2.0 * t.mid);
            sw = logaddexp(sw, (lX2 + t.tT + t.mid) - (lX + t.T));
            lX += t.T;
            lX2 += t.T2;
        })
    }
    double in3[3] = {sa, sb, sw}, ex3[3];
    cluster_scan_excl_k<OpLae, 3>(in3, sh, gpart + ev_gpart_stride(grp.nranks()), ex3, grp);
    const double Z0 = logaddexp(init.log_Z_mean, ex3[0]);
    const double dZ20 = logaddexp(init.log_dZ2_mean, ex3[1]);
    const double W0 = logaddexp(init.log_ZX_mean - init.log_X_mean, ex3[2]);
    // pass 3: Z2
    double sc = kNegInf;
    {
        double lX = X0, lX2 = X20, lW = W0;
        NSB_EV_LOOP({
            (void) i; (void) logL;
            const double zx_prev = lX + lW;
            sc = logaddexp(sc, logaddexp(kLog2 + zx_prev + t.t + t.mid, lX2 + t.t2 + 2.0 * t.mid));
            lW = logaddexp(lW, (lX2 + t.tT + t.mid) - (lX + t.T));
            lX += t.T;
            lX2 += t.T2;
        })
    }
    double in1[1] = {sc}, ex1[1];
    cluster_scan_excl_k<OpLae, 1>(in1, sh, gpart + 2 * ev_gpart_stride(grp.nranks()), ex1, grp);
    const double Z20 = logaddexp(init.log_Z2_mean, ex1[0]);
    // pass 4: outputs (skipped by threads that own neither a requested position nor per-sample rows)
    const bool wanted = out.per_sample || (out.mid && out.mark - 1 >= b && out.mark - 1 < e) || (out.fin && M - 1 >= b && M - 1 < e);
    if (wanted) {
        double lX = X0, lX2 = X20, lW = W0, lZ = Z0, ldZ2 = dZ20, lZ2 = Z20;
        NSB_EV_LOOP({
            const double dZ = lX + t.t + t.mid;
            const double x2t2m2 = lX2 + t.t2 + 2.0 * t.mid;
            const double zx_prev = lX + lW;
            lZ2 = logaddexp(logaddexp(lZ2, kLog2 + zx_prev + t.t + t.mid), x2t2m2);
            lZ = logaddexp(lZ, dZ);
            ldZ2 = logaddexp(ldZ2, x2t2m2);
            lW = logaddexp(lW, (lX2 + t.tT + t.mid) - (lX + t.T));
            lX += t.T;
            lX2 += t.T2;
            NsEvidenceCalc c;
            c.log_L = logL;
            c.log_X_mean = lX;
            c.log_X2_mean = lX2;
            c.log_Z_mean = lZ;
            c.log_ZX_mean = lX + lW;
            c.log_Z2_mean = lZ2;
            c.log_dZ_mean = dZ;
            c.log_dZ2_mean = ldZ2;
            if (out.per_sample) {
                double *p = out.per_sample + i;
                p[0 * M] = c.log_L;
                p[1 * M] = c.log_X_mean;
                p[2 * M] = c.log_X2_mean;
                p[3 * M] = c.log_Z_mean;
                p[4 * M] = c.log_ZX_mean;
                p[5 * M] = c.log_Z2_mean;
                p[6 * M] = c.log_dZ_mean;
                p[7 * M] = c.log_dZ2_mean;
            }
            if (out.mid && i == out.mark - 1) *out.mid = c;
            if (out.fin && i == M - 1) *out.fin = c;
        })
    }
#undef NSB_EV_LOOP
    if (M == 0 && gtid == 0) {
        if (out.fin) *out.fin = init;
        if (out.mid) *out.mid = init;
    } else if (out.mid && out.mark == 0 && gtid == 0) {
        *out.mid = init;
    }
}

constexpr int kEvCluster = 8;    // CTAs per cluster (portable maximum)
constexpr int kEvMaxGrid = 256;  // CTAs of the cooperative form (<= kEvThreads: the rank scan is one block scan)
constexpr int kEvThreads = 512;  // threads per CTA

__global__ void __cluster_dims__(kEvCluster, 1, 1) __launch_bounds__(kEvThreads)
k_evidence_stats(EvSeq q, NsEvidenceCalc init, EvOut out, double *gpart) {
    __shared__ double sh[3][34];
    ClusterSync grp;
    evidence_scan_block<0>(q, init, out, sh, gpart, grp);
}

// Long dead sets (final pass, M up to 1e7): up to one CTA per SM, launched cooperatively (co-residency is what
// the software barrier needs).  The scans are bound by the FP64 pipe (about 25 log / exp per element and pass),
// not by HBM: 80 M bytes of traffic against ~3000 FP64 instructions per element.
__global__ void __launch_bounds__(kEvThreads)
k_evidence_stats_grid(EvSeq q, NsEvidenceCalc init, EvOut out, double *gpart, unsigned *bar) {
    __shared__ double sh[3][34];
    GridSync grp{bar, 0u};
    evidence_scan_block<0>(q, init, out, sh, gpart, grp);
}

// =================================================================================================
// logsumexp (single CTA, two passes over the data).
// =================================================================================================
__global__ void __launch_bounds__(1024) k_logsumexp(const double *x, long long n, double *out) {
    __shared__ double sh[33];
    const double kNegInf = -__longlong_as_double(0x7FF0000000000000ll);
    double mx = kNegInf;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m2 = kNegInf;
        for (int w = 0; w < (blockDim.x >> 5); ++w) m2 = fmax(m2, sh[w]);
        sh[32] = m2;
    }
    __syncthreads();
    mx = sh[32];
    __syncthreads();
    const double shift = (mx == kNegInf || mx != mx || mx == -kNegInf) ? 0.0 : mx;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += exp(x[i] - shift);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
        out[0] = log(t) + shift;
    }
}

// =================================================================================================
// sample_evidence (/root/reference/src/jaxns/utils.py:433-476): S stochastic simulations of the shrinkage.
// One CTA per simulation s: key_s = split(key, S)[s]; element i draws log T_i = log(uniform(split(key_s, M)[i], ()))
// / n_i; log X = cumsum(log T) (block add-scan of per-thread chunk sums), dZ_i = (X_{i-1} - X_i) L_i with the
// LogSpace subtraction of internals/log_semiring.py:28-48, log Z = logsumexp_i dZ_i (block logaddexp reduction).
// The draws are recomputed in the second pass (2 Threefry blocks + one log each) instead of being stored:
// the job is integer-ALU bound, S * M * 16 bytes of scratch would make it HBM bound.
// =================================================================================================
__device__ __forceinline__ double sample_evidence_log_T(Key ks, long long i, const double *nlive) {
    const Key ki = split_child(ks, (uint64_t) i);
    return log(uniform01(ki, 0)) / nlive[i];
}

__global__ void __launch_bounds__(1024) k_sample_evidence(Key key, const double *nlive, const double *logL, long long M,
                                                          double *out) {
    __shared__ double sh[1][34];
    const Key ks = split_child(key, (uint64_t) blockIdx.x);
    const long long per = (M + blockDim.x - 1) / blockDim.x;
    const long long b = min(M, (long long) threadIdx.x * per), e = min(M, b + per);
    double sT = 0.0;
    for (long long i = b; i < e; ++i) sT += sample_evidence_log_T(ks, i, nlive);
    double in1[1] = {sT}, ex1[1];
    block_scan_excl_k<OpAdd, 1>(in1, sh, ex1);
    __syncthreads();
    double lX = ex1[0];  // log X before this thread's chunk (init log X = 0)
    double acc = OpLae::id();
    for (long long i = b; i < e; ++i) {
        const double nx = lX + sample_evidence_log_T(ks, i, nlive);
        const double delta = -fabs(nx - lX);
        // X - next_X = signed_logaddexp(log X, +, log next_X, -): NaN delta (inf - inf) -> sum of the logs
        const double diff = (delta != delta) ? lX + nx : fmax(lX, nx) + log1p(-exp(delta));
        acc = logaddexp(acc, diff + logL[i]);
        lX = nx;
    }
    double in2[1] = {acc}, ex2[1];
    block_scan_excl_k<OpLae, 1>(in2, sh, ex2);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0][32];  // CTA aggregate of the scan = logsumexp over all elements
}

}  // namespace nsb

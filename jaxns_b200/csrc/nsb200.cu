// C ABI of the B200-native nested-sampling hot path (see include/nsb200.h for the contract and the
// reference interfaces each entry point replaces).  Single translation unit: kernels are templates
// in the .cuh headers next to this file.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <chrono>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "ns_engine.cuh"
#include "ns_slice.cuh"
#include "ns_slice_mma.cuh"
#include "ns_split.cuh"

using namespace nsb;

// -------------------------------------------------------------------------------------------------
// error plumbing
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return 1;
}

#define NSB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define NSB_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" int nsb200_abi_version(void) { return NSB200_ABI_VERSION; }
extern "C" const char *nsb200_last_error(void) { return g_last_error.c_str(); }

// Tuning / A-B knobs.  Each is an integer option whose default comes from an environment variable read ONCE per
// process (getenv walks the whole environment: it has no place on a per-launch path); nsb200_set_option overrides
// it at run time (tests A/B kernels inside one process).  None of them changes results.
enum {
    OPT_SPEC, OPT_TPB, OPT_SLICE_MMA, OPT_MMA_P, OPT_MMA_WPB, OPT_MERGE_BRUTE, OPT_GEN_MODE, OPT_GEN_SMS, OPT_GEN_TPB,
    OPT_EPI_CLUSTER, OPT_DEPTH, OPT_TRACE, OPT_GEN_FENCE, OPT_SPECULATE, OPT_EPI_CTAS, OPT_EPI_PRIO, OPT_COUNT
};
struct NsOption {
    const char *name;
    int dflt;
    int value;
    bool loaded;
};
static NsOption g_opts[OPT_COUNT] = {
    {"NSB200_SPEC", 0, 0, false},        {"NSB200_TPB", 0, 0, false},         {"NSB200_SLICE_MMA", 0, 0, false},
    {"NSB200_MMA_P", 0, 0, false},       {"NSB200_MMA_WPB", 4, 0, false},     {"NSB200_MERGE_BRUTE", 0, 0, false},
    {"NSB200_GEN_MODE", 3, 0, false},    {"NSB200_GEN_SMS", 0, 0, false},     {"NSB200_GEN_TPB", 0, 0, false},
    {"NSB200_EPI_CLUSTER", 0, 0, false}, {"NSB200_DEPTH", 4, 0, false},       {"NSB200_TRACE", 0, 0, false},
    {"NSB200_GEN_FENCE", 0, 0, false},  {"NSB200_SPECULATE", 2, 0, false},  {"NSB200_EPI_CTAS", 0, 0, false},
    {"NSB200_EPI_PRIO", 1, 0, false},
};
static int opt(int id) {
    NsOption &o = g_opts[id];
    if (!o.loaded) {
        const char *e = getenv(o.name);
        o.value = (e && *e) ? atoi(e) : o.dflt;
        o.loaded = true;
    }
    return o.value;
}

extern "C" int nsb200_set_option(const char *name, int32_t value) {
    if (!name) return fail("option name is NULL");
    for (int i = 0; i < OPT_COUNT; ++i) {
        if (strcmp(name, g_opts[i].name) == 0) {
            if (value < 0) {
                g_opts[i].loaded = false;  // back to the environment / built-in default
            } else {
                g_opts[i].value = value;
                g_opts[i].loaded = true;
            }
            return 0;
        }
    }
    return fail("unknown option %s", name);
}

// -------------------------------------------------------------------------------------------------
// launch geometry: group size G (lanes per chain) and DPL (dims per lane)
// -------------------------------------------------------------------------------------------------
struct Geometry {
    int G, DPL, DP;
};

static int pick_geometry(int D, Geometry &g) {
    if (D < 1) return fail("model.D must be >= 1, got %d", D);
    if (D > 256) return fail("model.D = %d exceeds the supported maximum of 256", D);
    int G = 1;
    while (G < D && G < 32) G <<= 1;
    int dpl = (D + G - 1) / G;
    int DPL = 1;
    while (DPL < dpl) DPL <<= 1;
    g.G = G;
    g.DPL = DPL;
    g.DP = G * DPL;
    return 0;
}

static int check_model(const NsModelDesc *m, bool allow_external = false) {
    if (!m) return fail("model is NULL");
    if (m->family == NSB200_FAM_EXTERNAL && !allow_external)
        return fail("family EXTERNAL is evaluated by the caller: use the nsb200_split_* / nsb200_engine_split_* entry points");
    if (m->family < 0 || m->family > NSB200_FAM_EXTERNAL) return fail("unknown likelihood family %d", m->family);
    if (m->prior_kind != NSB200_PRIOR_UNIFORM && m->prior_kind != NSB200_PRIOR_NORMAL)
        return fail("unknown prior kind %d", m->prior_kind);
    if (!m->prior_a || !m->prior_b) return fail("model prior arrays are NULL");
    const long long D = m->D, K = m->K;
    long long need = 0;
    switch (m->family) {
        case NSB200_FAM_GAUSS_DENSE: need = 1 + D + D * D; break;
        case NSB200_FAM_GAUSS_MIX_DIAG: need = K * (1 + 2 * D); break;
        case NSB200_FAM_SHELLS: need = K * (2 + D); break;
        default: need = 0;
    }
    if ((m->family == NSB200_FAM_GAUSS_MIX_DIAG || m->family == NSB200_FAM_SHELLS) && K < 1)
        return fail("family %d needs K >= 1", m->family);
    if (need > 0 && (!m->params || m->n_params < need))
        return fail("family %d with D=%lld K=%lld needs %lld params, got %lld", m->family, D, K, need,
                    (long long) m->n_params);
    return 0;
}

template <typename KernelT>
static int set_smem(KernelT kernel, size_t bytes) {
    if (bytes > 227 * 1024) return fail("model needs %zu bytes of shared memory per CTA (> 227 KB)", bytes);
    if (bytes > 48 * 1024) NSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return 0;
}

// Compile-time (G, DPL) instantiations: G < 32 only with DPL == 1.
// -DNSB_FAST_BUILD (kernel experiments): only the config-2 geometry (G = 32, DPL = 1), P = 1, dense Gaussian.
#ifdef NSB_FAST_BUILD
#define NSB_DISPATCH_GEOM(GEOM, ...)                                                \
    if ((GEOM).G == 32 && (GEOM).DPL == 1) {                                        \
        constexpr int kG = 32, kDPL = 1;                                            \
        __VA_ARGS__;                                                                \
    } else {                                                                        \
        return fail("NSB_FAST_BUILD supports only 17 <= D <= 32");                  \
    }
#else
#define NSB_DISPATCH_GEOM(GEOM, ...)                                                \
    if ((GEOM).G == 32) {                                                           \
        switch ((GEOM).DPL) {                                                       \
            case 1: { constexpr int kG = 32, kDPL = 1; __VA_ARGS__; } break;        \
            case 2: { constexpr int kG = 32, kDPL = 2; __VA_ARGS__; } break;        \
            case 4: { constexpr int kG = 32, kDPL = 4; __VA_ARGS__; } break;        \
            case 8: { constexpr int kG = 32, kDPL = 8; __VA_ARGS__; } break;        \
            default: return fail("unsupported DPL %d", (GEOM).DPL);                 \
        }                                                                           \
    } else {                                                                        \
        switch ((GEOM).G) {                                                         \
            case 1: { constexpr int kG = 1, kDPL = 1; __VA_ARGS__; } break;         \
            case 2: { constexpr int kG = 2, kDPL = 1; __VA_ARGS__; } break;         \
            case 4: { constexpr int kG = 4, kDPL = 1; __VA_ARGS__; } break;         \
            case 8: { constexpr int kG = 8, kDPL = 1; __VA_ARGS__; } break;         \
            case 16: { constexpr int kG = 16, kDPL = 1; __VA_ARGS__; } break;       \
            default: return fail("unsupported group size %d", (GEOM).G);            \
        }                                                                           \
    }
#endif

// Number of proposals evaluated speculatively per shrink round (ns_slice.cuh).  NSB200_SPEC
// overrides it for the D <= 32 instantiation (tuning knob; results do not depend on it).
static int pick_spec(const Geometry &g) {
    if (g.G == 32 && g.DPL == 1) {
        const int spec = opt(OPT_SPEC);
        if (spec == 1 || spec == 2 || spec == 4) return spec;
        // measured on config 2 (profiles/r1/spec_sweep_r1.txt): P = 2 with the batched quantile is 1.4 % faster than
        // P = 1 end to end (rounds per slice 4.8 -> 2.6, instructions +5 %), P = 4 is 17 % slower
        return 2;
    }
    return 2;
}

extern "C" int nsb200_read_key(const uint32_t *device_key, uint32_t out[2], nsb200_stream_t stream) {
    if (!device_key || !out) return fail("NULL argument");
    NSB_CUDA(cudaMemcpyAsync(out, device_key, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    NSB_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}

// -------------------------------------------------------------------------------------------------
// jax.random primitives
// -------------------------------------------------------------------------------------------------
__global__ void k_threefry(Key key, const uint32_t *x0, const uint32_t *x1, long long n, uint32_t *o0, uint32_t *o1) {
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = x0[i], b = x1[i];
    threefry2x32(key.a, key.b, a, b);
    o0[i] = a;
    o1[i] = b;
}

// mode 0: split -> out u32[n,2]; 1: bits64; 2: uniform(lo,hi); 3: normal
__global__ void k_random(Key key, long long n, int mode, double lo, double hi, void *out) {
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) {
        Key c = split_child(key, (uint64_t) i);
        ((uint32_t *) out)[2 * i] = c.a;
        ((uint32_t *) out)[2 * i + 1] = c.b;
    } else if (mode == 1) {
        ((uint64_t *) out)[i] = bits64(key, (uint64_t) i);
    } else if (mode == 2) {
        ((double *) out)[i] = uniform_lohi(bits64(key, (uint64_t) i), lo, hi);
    } else {
        ((double *) out)[i] = normal_from_bits(bits64(key, (uint64_t) i));
    }
}

static inline int grid_for(long long n, int threads) { return (int) ((n + threads - 1) / threads); }

extern "C" int nsb200_threefry2x32(const uint32_t key[2], const uint32_t *x0, const uint32_t *x1, int64_t n,
                                   uint32_t *out0, uint32_t *out1, nsb200_stream_t stream) {
    if (n <= 0) return 0;
    k_threefry<<<grid_for(n, 256), 256, 0, (cudaStream_t) stream>>>(Key{key[0], key[1]}, x0, x1, n, out0, out1);
    NSB_LAUNCH_CHECK();
    return 0;
}

static int launch_random(const uint32_t key[2], int64_t n, int mode, double lo, double hi, void *out, nsb200_stream_t stream) {
    if (n <= 0) return 0;
    if (!out) return fail("output pointer is NULL");
    k_random<<<grid_for(n, 256), 256, 0, (cudaStream_t) stream>>>(Key{key[0], key[1]}, n, mode, lo, hi, out);
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_random_split(const uint32_t key[2], int64_t n, uint32_t *out, nsb200_stream_t stream) {
    return launch_random(key, n, 0, 0, 0, out, stream);
}
extern "C" int nsb200_random_bits64(const uint32_t key[2], int64_t n, uint64_t *out, nsb200_stream_t stream) {
    return launch_random(key, n, 1, 0, 0, out, stream);
}
extern "C" int nsb200_random_uniform(const uint32_t key[2], int64_t n, double lo, double hi, double *out,
                                     nsb200_stream_t stream) {
    return launch_random(key, n, 2, lo, hi, out, stream);
}
extern "C" int nsb200_random_normal(const uint32_t key[2], int64_t n, double *out, nsb200_stream_t stream) {
    return launch_random(key, n, 3, 0, 0, out, stream);
}

// -------------------------------------------------------------------------------------------------
// model / samplers
// -------------------------------------------------------------------------------------------------
static size_t sampler_smem_bytes(const NsModelDesc &m, const Geometry &g, int P, bool slice) {
    const size_t per_chain = chain_smem_doubles(g.G, g.DPL, P, slice);
    return 8 * (model_smem_doubles(m.family, m.D, g.G, g.DPL, m.K) + (kThreadsPerBlock / g.G) * per_chain);
}

extern "C" int nsb200_forward_batch(const NsModelDesc *model, const double *U, int64_t n, double *out_logL,
                                    double *out_X, nsb200_stream_t stream) {
    if (check_model(model)) return 1;
    if (n <= 0) return 0;
    if (!U) return fail("U is NULL");
    Geometry g;
    if (pick_geometry(model->D, g)) return 1;
    ForwardArgs a{*model, U, out_logL, out_X, n};
    const size_t smem = sampler_smem_bytes(*model, g, 1, false);
    const int per_block = kThreadsPerBlock / g.G;
    NSB_DISPATCH_GEOM(g, {
        if (set_smem(k_forward<kG, kDPL>, smem)) return 1;
        k_forward<kG, kDPL><<<grid_for(n, per_block), kThreadsPerBlock, smem, (cudaStream_t) stream>>>(a);
    });
    NSB_LAUNCH_CHECK();
    return 0;
}

// c[q] = logaddexp-cumsum of q + 1 zeros, continued from c[start - 1] (start = 0: from -inf).  The recurrence is
// serial by definition (it has to round like the reference's sequential cumulative_logsumexp, log_semiring.py:51-92).
__global__ void k_seed_table(long long start, long long N, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double acc = start > 0 ? out[start - 1] : -__longlong_as_double(0x7FF0000000000000ll);
    for (long long q = start; q < N; ++q) {
        acc = logaddexp(acc, 0.0);
        out[q] = acc;
    }
}

// The table for N is a prefix of the table for any N' > N and depends on nothing else, so the library keeps ONE per
// process and device and only ever computes the missing tail (0.36 us per entry on one thread: 1.2 ms at N = 3200,
// 36 ms at N = 1e5 -- paid once instead of by every new engine).
struct SeedTableCache {
    double *ptr = nullptr;
    long long n = 0;
    int device = -1;
    cudaEvent_t ready = nullptr;
};
static SeedTableCache g_seed_cache;
static std::mutex g_seed_mutex;

// Enqueues on `st` a copy of the first N table entries into `out` (device), extending the cache first if needed.
static int seed_table_into(long long N, double *out, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_seed_mutex);
    int dev = -1;
    NSB_CUDA(cudaGetDevice(&dev));
    SeedTableCache &c = g_seed_cache;
    if (c.device != dev) {  // one device per process in this framework: a device switch simply starts over
        if (c.ptr) cudaFree(c.ptr);
        if (c.ready) cudaEventDestroy(c.ready);
        c = SeedTableCache();
        c.device = dev;
        NSB_CUDA(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    }
    if (N > c.n) {
        const long long n_new = N > 2 * c.n ? N : 2 * c.n;
        double *q = nullptr;
        NSB_CUDA(cudaMalloc(&q, (size_t) n_new * 8));
        if (c.n > 0) {
            NSB_CUDA(cudaStreamWaitEvent(st, c.ready, 0));
            NSB_CUDA(cudaMemcpyAsync(q, c.ptr, (size_t) c.n * 8, cudaMemcpyDeviceToDevice, st));
        }
        k_seed_table<<<1, 1, 0, st>>>(c.n, n_new, q);
        NSB_LAUNCH_CHECK();
        NSB_CUDA(cudaEventRecord(c.ready, st));
        if (c.ptr) {
            NSB_CUDA(cudaDeviceSynchronize());  // the old table may still be read by copies enqueued on other streams
            cudaFree(c.ptr);
        }
        c.ptr = q;
        c.n = n_new;
    } else {
        NSB_CUDA(cudaStreamWaitEvent(st, c.ready, 0));
    }
    NSB_CUDA(cudaMemcpyAsync(out, c.ptr, (size_t) N * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int nsb200_seed_table(int64_t N, double *out, nsb200_stream_t stream) {
    if (N <= 0) return 0;
    if (!out) return fail("out is NULL");
    return seed_table_into(N, out, (cudaStream_t) stream);
}

static int launch_draw(const NsModelDesc *model, Key key, const double *contour, long long begin, long long end,
                       double *out_U, double *out_logL, long long *out_nevals, int uniform_sampler,
                       cudaStream_t st) {
    if (check_model(model)) return 1;
    if (end <= begin) return 0;
    Geometry g;
    if (pick_geometry(model->D, g)) return 1;
    DrawArgs a{*model, key, contour, out_U, out_logL, out_nevals, begin, end, uniform_sampler};
    const size_t smem = sampler_smem_bytes(*model, g, 1, false);
    const int per_block = kThreadsPerBlock / g.G;
    NSB_DISPATCH_GEOM(g, {
        if (set_smem(k_draw<kG, kDPL>, smem)) return 1;
        k_draw<kG, kDPL><<<grid_for(end - begin, per_block), kThreadsPerBlock, smem, st>>>(a);
    });
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_init_batch(const NsModelDesc *model, const uint32_t sample_key[2], int64_t N, int64_t begin,
                                 int64_t end, double *out_U, double *out_logL, int64_t *out_nevals,
                                 nsb200_stream_t stream) {
    if (begin < 0 || end > N || begin > end) return fail("bad range [%lld, %lld) of %lld", (long long) begin, (long long) end, (long long) N);
    if (!out_U || !out_logL || !out_nevals) return fail("output pointer is NULL");
    return launch_draw(model, Key{sample_key[0], sample_key[1]}, nullptr, begin, end, out_U, out_logL,
                       (long long *) out_nevals, 0, (cudaStream_t) stream);
}

extern "C" int nsb200_uniform_batch(const NsModelDesc *model, const uint32_t key[2], const double *contour,
                                    int64_t num_samples, int64_t chain_begin, int64_t chain_end, double *out_U,
                                    double *out_logL, int64_t *out_nevals, nsb200_stream_t stream) {
    if (chain_begin < 0 || chain_end > num_samples || chain_begin > chain_end) return fail("bad chain range");
    if (!contour) return fail("contour is NULL");
    if (!out_U || !out_logL || !out_nevals) return fail("output pointer is NULL");
    return launch_draw(model, Key{key[0], key[1]}, contour, chain_begin, chain_end, out_U, out_logL,
                       (long long *) out_nevals, 1, (cudaStream_t) stream);
}

// CTA size of the slice kernel.  When the factor lives in registers (D <= 32) nothing big is staged
// per CTA, so one chain per CTA gives the block scheduler the finest grain to balance the 148 SMs.
static int slice_threads(const Geometry &g) {
    int tpb = (g.G == 32 && g.DPL == 1) ? 32 : kThreadsPerBlock;
    const int v = opt(OPT_TPB);
    if (v == 32 || v == 64 || v == 128) tpb = v;
    return tpb < g.G ? g.G : tpb;
}

// pdl: programmatic dependent launch -- the kernel may start as soon as every CTA of the PREVIOUS kernel in
// the stream has executed griddepcontrol.launch_dependents (or exited), instead of after its completion.
template <int G, int DPL, int P>
static int launch_slice_t(const SliceArgs &a, const Geometry &g, cudaStream_t st, bool pdl) {
    const long long n = a.chain_end - a.chain_begin;
    const int tpb = slice_threads(g);
    const size_t per_chain = chain_smem_doubles(g.G, g.DPL, P, true);
    const size_t smem = 8 * (model_smem_doubles(a.model.family, a.model.D, g.G, g.DPL, a.model.K) + (tpb / g.G) * per_chain);
    if (set_smem(k_slice_chains<G, DPL, P>, smem)) return 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned) grid_for(n, tpb / G));
    cfg.blockDim = dim3((unsigned) tpb);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    NSB_CUDA(cudaLaunchKernelEx(&cfg, k_slice_chains<G, DPL, P>, a));
    return 0;
}

// ---- FP64 tensor-core slice kernel (ns_slice_mma.cuh): dense Gaussian, D <= 32, pre-generated streams ----------
// Which slice kernel runs the dense Gaussian family with D <= 32 (NSB200_SLICE_MMA): 0 (default) = lane-per-dimension
// kernel, 1 = DMMA kernel, 2 / 4 = warp teams, 8 = DMMA from kSliceMmaMinChains chains per GPU.  The DMMA kernel is a
// THROUGHPUT design: a warp carries 8 proposal columns and issues ~1150 instructions per round with little
// instruction-level parallelism in its bookkeeping, so a round takes ~5400 cycles against ~2400 for the lane
// kernel's one-chain warp (profiles/r2/mma_cycles_r2.txt).  A nested-sampling launch is S sequential slices per
// chain, so with few chains the launch time is rounds x round latency (config 2: 0.55 ms lane, 0.77 ms DMMA); with
// many chains the lane kernel's 2.7x more warps hide the same latencies, and the two meet only at 25600 chains per
// GPU (8.07 vs 8.18 ms, profiles/r2/slice_crossover_r2.txt).  Measured, not assumed: the lane kernel is the default
// at every size, the DMMA kernel stays selectable and parity-tested.
constexpr long long kSliceMmaMinChains = 32768;

static bool slice_mma_eligible(const SliceArgs &a) {
    const int mode = opt(OPT_SLICE_MMA);
    const bool by_size = mode == 8 && (a.chain_end - a.chain_begin) >= kSliceMmaMinChains;
    return (mode == 1 || by_size) && a.model.family == NSB200_FAM_GAUSS_DENSE && a.model.D <= 32 && a.pre_dirs != nullptr;
}

// Speculative proposals per chain and round: a warp carries 8 proposal columns = 8 / P chains.  One warp per SM
// sub-partition keeps the FP64 pipe of that sub-partition busy on its own (8 independent quantiles per lane), so
// P grows as the chains get fewer (strong scaling over GPUs) until the warps no longer cover the sub-partitions.
static int slice_mma_spec(long long n_chains) {
    const int forced = opt(OPT_MMA_P);
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    static int smsp = 0;
    if (!smsp) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        smsp = 4 * sms;
    }
    if (n_chains * 4 / 8 <= smsp) return 4;
    if (n_chains * 2 / 8 <= smsp) return 2;
    return 1;
}

template <int NB, int P>
static int launch_slice_mma_t(const SliceArgs &a, cudaStream_t st, bool pdl) {
    const int wpb_env = opt(OPT_MMA_WPB);
    const int wpb = (wpb_env >= 1 && wpb_env <= 4) ? wpb_env : 4;
    const long long n = a.chain_end - a.chain_begin;
    const long long per_cta = (long long) wpb * (8 / P);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned) ((n + per_cta - 1) / per_cta));
    cfg.blockDim = dim3((unsigned) (32 * wpb));
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    NSB_CUDA(cudaLaunchKernelEx(&cfg, k_slice_chains_mma<NB, P>, a));
    return 0;
}

template <int NB>
static int launch_slice_mma_nb(const SliceArgs &a, cudaStream_t st, bool pdl) {
    switch (slice_mma_spec(a.chain_end - a.chain_begin)) {
        case 1: return launch_slice_mma_t<NB, 1>(a, st, pdl);
        case 4: return launch_slice_mma_t<NB, 4>(a, st, pdl);
        default: return launch_slice_mma_t<NB, 2>(a, st, pdl);
    }
}

static int launch_slice_mma(const SliceArgs &a, cudaStream_t st, bool pdl) {
    int rc;
    const int D = a.model.D;
    if (D <= 8) rc = launch_slice_mma_nb<1>(a, st, pdl);
    else if (D <= 16) rc = launch_slice_mma_nb<2>(a, st, pdl);
    else if (D <= 24) rc = launch_slice_mma_nb<3>(a, st, pdl);
    else rc = launch_slice_mma_nb<4>(a, st, pdl);
    if (rc) return rc;
    NSB_LAUNCH_CHECK();
    return 0;
}

// ---- warp teams (ns_slice.cuh k_slice_chains_team): W warps per chain, one speculative proposal each ------------
template <int W>
static int launch_slice_team_t(const SliceArgs &a, cudaStream_t st, bool pdl) {
    const long long n = a.chain_end - a.chain_begin;
    const size_t smem = 8 * (model_smem_doubles(a.model.family, a.model.D, 32, 1, a.model.K, true) +
                             (size_t) W * chain_smem_doubles(32, 1, 1, true) + 2 * W);
    if (set_smem(k_slice_chains_team<W>, smem)) return 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned) n);
    cfg.blockDim = dim3((unsigned) (32 * W));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    NSB_CUDA(cudaLaunchKernelEx(&cfg, k_slice_chains_team<W>, a));
    NSB_LAUNCH_CHECK();
    return 0;
}

static int launch_slice(const SliceArgs &a, cudaStream_t st, bool pdl = false) {
    Geometry g;
    if (pick_geometry(a.model.D, g)) return 1;
    if (a.chain_end <= a.chain_begin) return 0;
    if ((opt(OPT_SLICE_MMA) == 2 || opt(OPT_SLICE_MMA) == 4) && a.model.family == NSB200_FAM_GAUSS_DENSE && g.G == 32 && g.DPL == 1 && a.pre_dirs) {
        return opt(OPT_SLICE_MMA) == 4 ? launch_slice_team_t<4>(a, st, pdl) : launch_slice_team_t<2>(a, st, pdl);
    }
    if (slice_mma_eligible(a)) return launch_slice_mma(a, st, pdl);
    const int P = pick_spec(g);
    if (g.G == 32 && g.DPL == 1) {
        int rc;
#ifdef NSB_FAST_BUILD_P1
        rc = launch_slice_t<32, 1, 1>(a, g, st, pdl);
#else
        if (P == 1) rc = launch_slice_t<32, 1, 1>(a, g, st, pdl);
        else if (P == 4) rc = launch_slice_t<32, 1, 4>(a, g, st, pdl);
        else rc = launch_slice_t<32, 1, 2>(a, g, st, pdl);
#endif
        if (rc) return rc;
    } else {
#ifdef NSB_FAST_BUILD
        return fail("NSB_FAST_BUILD supports only 17 <= D <= 32");
#else
        NSB_DISPATCH_GEOM(g, { if (launch_slice_t<kG, kDPL, 2>(a, g, st, pdl)) return 1; });
#endif
    }
    NSB_LAUNCH_CHECK();
    return 0;
}

static int slice_batch_args(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                            const double *contour, const double *live_U, const double *live_logL,
                            const double *seed_table, double *out_U, double *out_logL, int64_t *out_nevals,
                            double *ph_U, double *ph_logL, SliceArgs &a) {
    if (check_model(model)) return 1;
    if (!p) return fail("params is NULL");
    if (p->num_slices < 1) return fail("num_slices should be >= 1, got %d", p->num_slices);
    if (p->num_phantom < 0) return fail("num_phantom_save should be >= 0, got %d", p->num_phantom);
    if (p->num_phantom >= p->num_slices)
        return fail("num_phantom_save should be < num_slices, got %d >= %d", p->num_phantom, p->num_slices);
    if (p->num_live < 1) return fail("num_live must be >= 1");
    if (p->chain_begin < 0 || p->chain_end > p->num_samples || p->chain_begin > p->chain_end)
        return fail("bad chain range [%lld, %lld) of %lld", (long long) p->chain_begin, (long long) p->chain_end,
                    (long long) p->num_samples);
    if (!key || !contour || !live_U || !live_logL || !seed_table) return fail("input pointer is NULL");
    if (!out_U || !out_logL || !out_nevals) return fail("output pointer is NULL");
    if (p->num_phantom > 0 && (!ph_U || !ph_logL)) return fail("phantom outputs are NULL but num_phantom > 0");
    memset(&a, 0, sizeof(a));
    a.model = *model;
    a.key = Key{key[0], key[1]};
    a.contour = contour;
    a.live_U = live_U;
    a.live_logL = live_logL;
    a.seed_table = seed_table;
    a.out_U = out_U;
    a.out_logL = out_logL;
    a.out_nevals = (long long *) out_nevals;
    a.ph_U = ph_U;
    a.ph_logL = ph_logL;
    a.N = p->num_live;
    a.chain_begin = p->chain_begin;
    a.chain_end = p->chain_end;
    a.S = p->num_slices;
    a.k = p->num_phantom;
    a.midpoint = p->midpoint_shrink;
    return 0;
}

extern "C" int nsb200_slice_batch(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                                  const double *contour, const double *live_U, const double *live_logL,
                                  const double *seed_table, double *out_U, double *out_logL, int64_t *out_nevals,
                                  double *ph_U, double *ph_logL, nsb200_stream_t stream) {
    SliceArgs a;
    if (slice_batch_args(model, p, key, contour, live_U, live_logL, seed_table, out_U, out_logL, out_nevals, ph_U,
                         ph_logL, a))
        return 1;
    return launch_slice(a, (cudaStream_t) stream);
}

// chain streams of n chains x S slices: directions [n][S][D], proposal uniforms [n][S][kPre], continuation keys
// [n][S], + one error word
static size_t slice_streams_bytes(int D, int S, long long n) {
    const size_t rows = (size_t) n * (size_t) S;
    return align256(rows * D * 8) + align256(rows * kPre * 8) + align256(rows * 8) + 256;
}

extern "C" int64_t nsb200_slice_streams_bytes(int32_t D, int32_t num_slices, int64_t n_chains) {
    if (D < 1 || num_slices < 1 || n_chains < 0) return -1;
    return (int64_t) slice_streams_bytes(D, num_slices, n_chains);
}

static int launch_chain_streams(const StreamArgs &sa, int ctas, int tpb, size_t smem, cudaStream_t st) {
    const long long warps = (sa.chain_end - sa.chain_begin) * ((sa.S + 31) / 32);
    const long long wpc = tpb / 32;
    if ((long long) ctas * wpc > warps) ctas = (int) ((warps + wpc - 1) / wpc);
    if (ctas < 1) return 0;
    if (sa.D <= 32) k_chain_streams<1><<<ctas, tpb, smem, st>>>(sa);
    else if (sa.D <= 128) k_chain_streams<4><<<ctas, tpb, smem, st>>>(sa);
    else k_chain_streams<8><<<ctas, tpb, smem, st>>>(sa);
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_slice_batch_ws(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                                     const double *contour, const double *live_U, const double *live_logL,
                                     const double *seed_table, double *out_U, double *out_logL, int64_t *out_nevals,
                                     double *ph_U, double *ph_logL, void *workspace, int64_t workspace_bytes,
                                     int32_t *error_flags, nsb200_stream_t stream) {
    SliceArgs a;
    if (slice_batch_args(model, p, key, contour, live_U, live_logL, seed_table, out_U, out_logL, out_nevals, ph_U,
                         ph_logL, a))
        return 1;
    const long long n = a.chain_end - a.chain_begin;
    if (n <= 0) return 0;
    Geometry g;
    if (pick_geometry(model->D, g)) return 1;
    cudaStream_t st = (cudaStream_t) stream;
    a.err = (int *) error_flags;  // optional DEVICE word, OR-ed with NSB200_ERR_* bits
    if (g.G >= 8) {  // narrower groups keep their in-kernel lane-parallel stream precompute
        if (!workspace) return fail("workspace is NULL");
        if (workspace_bytes < (int64_t) slice_streams_bytes(model->D, a.S, n))
            return fail("workspace too small: %lld < %zu (nsb200_slice_streams_bytes)", (long long) workspace_bytes,
                        slice_streams_bytes(model->D, a.S, n));
        const size_t rows = (size_t) n * (size_t) a.S;
        char *w = (char *) align256((size_t) workspace);
        StreamArgs sa;
        sa.key = a.key;
        sa.ctl = nullptr;
        sa.key_slot = 0;
        sa.chain_begin = a.chain_begin;
        sa.chain_end = a.chain_end;
        sa.S = a.S;
        sa.D = model->D;
        sa.dirs = (double *) w;
        w += align256(rows * model->D * 8);
        sa.us = (double *) w;
        w += align256(rows * kPre * 8);
        sa.rkeys = (uint2 *) w;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        if (launch_chain_streams(sa, sms, 1024, 0, st)) return 1;
        a.pre_dirs = sa.dirs;
        a.pre_us = sa.us;
        a.pre_rkeys = sa.rkeys;
    }
    return launch_slice(a, st);
}

// -------------------------------------------------------------------------------------------------
// split slice step around a caller-evaluated likelihood (ns_split.cuh)
// -------------------------------------------------------------------------------------------------
// The split kernels only use the prior part of the model: strip the likelihood so that no
// family-specific shared memory is staged.
static NsModelDesc prior_only(const NsModelDesc &m) {
    NsModelDesc p = m;
    p.family = NSB200_FAM_EXTERNAL;
    p.K = 0;
    p.params = nullptr;
    p.n_params = 0;
    return p;
}

static int launch_split_step(const SplitArgs &a, int mode, cudaStream_t st) {
    Geometry g;
    if (pick_geometry(a.model.D, g)) return 1;
    const long long n = a.chain_end - a.chain_begin;
    if (n <= 0) return 0;
    const size_t smem = 8 * model_smem_doubles(a.model.family, a.model.D, g.G, g.DPL, 0);
    const int per_block = kThreadsPerBlock / g.G;
    NSB_DISPATCH_GEOM(g, {
        if (set_smem(k_split_step<kG, kDPL>, smem)) return 1;
        k_split_step<kG, kDPL><<<grid_for(n, per_block), kThreadsPerBlock, smem, st>>>(a, mode);
    });
    NSB_LAUNCH_CHECK();
    return 0;
}

static int check_slice_params(const NsSliceParams *p) {
    if (!p) return fail("params is NULL");
    if (p->num_slices < 1) return fail("num_slices should be >= 1, got %d", p->num_slices);
    if (p->num_phantom < 0) return fail("num_phantom_save should be >= 0, got %d", p->num_phantom);
    if (p->num_phantom >= p->num_slices)
        return fail("num_phantom_save should be < num_slices, got %d >= %d", p->num_phantom, p->num_slices);
    if (p->chain_begin < 0 || p->chain_end > p->num_samples || p->chain_begin > p->chain_end)
        return fail("bad chain range [%lld, %lld) of %lld", (long long) p->chain_begin, (long long) p->chain_end,
                    (long long) p->num_samples);
    return 0;
}

extern "C" int64_t nsb200_split_workspace_bytes(int32_t D, int64_t n_chains, int32_t num_phantom) {
    if (D < 1 || n_chains < 0 || num_phantom < 0) return -1;
    return (int64_t) split_workspace_bytes(D, n_chains, num_phantom);
}

static int split_args(const NsModelDesc *model, const NsSliceParams *p, void *workspace, int64_t workspace_bytes,
                      SplitArgs &a) {
    if (check_model(model, true)) return 1;
    if (check_slice_params(p)) return 1;
    if (!workspace) return fail("workspace is NULL");
    const long long n = p->chain_end - p->chain_begin;
    if (workspace_bytes < (int64_t) split_workspace_bytes(model->D, n, p->num_phantom))
        return fail("workspace too small: %lld < %zu", (long long) workspace_bytes,
                    split_workspace_bytes(model->D, n, p->num_phantom));
    memset(&a, 0, sizeof(a));
    a.model = prior_only(*model);
    a.N = p->num_live;
    a.chain_begin = p->chain_begin;
    a.chain_end = p->chain_end;
    a.S = p->num_slices;
    a.k = p->num_phantom;
    a.midpoint = p->midpoint_shrink;
    a.grad_flags = p->split_flags & 3;
    a.P = (p->split_flags >> 8) & 0xF;
    if (a.P < 1) a.P = 1;
    if (a.P > kSplitMaxP) return fail("NsSliceParams.split_flags: at most %d proposals per round", kSplitMaxP);
    a.state = split_state_view(workspace, model->D, n, p->num_phantom);
    return 0;
}

extern "C" int nsb200_split_begin(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                                  const double *contour, const double *live_U, const double *live_logL,
                                  const double *seed_table, void *workspace, int64_t workspace_bytes, double *prop_U,
                                  double *prop_X, nsb200_stream_t stream) {
    SplitArgs a;
    if (split_args(model, p, workspace, workspace_bytes, a)) return 1;
    if (p->num_live < 1) return fail("num_live must be >= 1");
    if (!key || !contour || !live_U || !live_logL || !seed_table) return fail("input pointer is NULL");
    if (!prop_U) return fail("prop_U is NULL");
    a.key = Key{key[0], key[1]};
    a.contour = contour;
    a.live_U = live_U;
    a.live_logL = live_logL;
    a.seed_table = seed_table;
    a.prop_U = prop_U;
    a.prop_X = prop_X;
    return launch_split_step(a, 0, (cudaStream_t) stream);
}

extern "C" int nsb200_split_accept(const NsModelDesc *model, const NsSliceParams *p, const double *contour,
                                   const double *prop_logL, void *workspace, int64_t workspace_bytes, double *prop_U,
                                   double *prop_X, uint64_t *n_active, nsb200_stream_t stream) {
    SplitArgs a;
    if (split_args(model, p, workspace, workspace_bytes, a)) return 1;
    if (!contour || !prop_logL) return fail("input pointer is NULL");
    if (!prop_U) return fail("prop_U is NULL");
    a.contour = contour;
    a.prop_logL = prop_logL;
    a.prop_U = prop_U;
    a.prop_X = prop_X;
    a.active = (unsigned long long *) n_active;
    return launch_split_step(a, 1, (cudaStream_t) stream);
}

extern "C" int nsb200_split_grad_points(const NsModelDesc *model, const NsSliceParams *p, void *workspace,
                                        int64_t workspace_bytes, double *out_U, nsb200_stream_t stream) {
    SplitArgs a;
    if (split_args(model, p, workspace, workspace_bytes, a)) return 1;
    if (!out_U) return fail("out_U is NULL");
    const long long n = p->chain_end - p->chain_begin;
    if (n <= 0) return 0;
    k_split_export_U0<<<296, 256, 0, (cudaStream_t) stream>>>(nullptr, a.state, n, model->D, out_U);
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_split_grad_begin(const NsModelDesc *model, const NsSliceParams *p, const double *contour,
                                       const double *grad, void *workspace, int64_t workspace_bytes, double *prop_U,
                                       double *prop_X, uint64_t *n_active, nsb200_stream_t stream) {
    SplitArgs a;
    if (split_args(model, p, workspace, workspace_bytes, a)) return 1;
    if (!a.grad_flags) return fail("nsb200_split_grad_begin needs gradient bits in NsSliceParams.split_flags");
    if (!contour || !grad || !prop_U) return fail("NULL pointer");
    a.contour = contour;
    a.grad = grad;
    a.prop_U = prop_U;
    a.prop_X = prop_X;
    a.active = (unsigned long long *) n_active;
    return launch_split_step(a, 2, (cudaStream_t) stream);
}

extern "C" int nsb200_split_finish(const NsModelDesc *model, const NsSliceParams *p, void *workspace,
                                   int64_t workspace_bytes, double *out_U, double *out_logL, int64_t *out_nevals,
                                   double *ph_U, double *ph_logL, nsb200_stream_t stream) {
    SplitArgs a;
    if (split_args(model, p, workspace, workspace_bytes, a)) return 1;
    if (!out_U || !out_logL || !out_nevals) return fail("output pointer is NULL");
    if (p->num_phantom > 0 && (!ph_U || !ph_logL)) return fail("phantom outputs are NULL but num_phantom > 0");
    const long long n = p->chain_end - p->chain_begin;
    if (n <= 0) return 0;
    k_split_finish<<<296, 256, 0, (cudaStream_t) stream>>>(nullptr, a.state, n, model->D, p->num_phantom, out_U, out_logL,
                                                            (long long *) out_nevals, ph_U, ph_logL, nullptr, 0);
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_init_propose(const NsModelDesc *model, const uint32_t sample_key[2], int64_t N, int64_t begin,
                                   int64_t end, int32_t round, const uint8_t *need, double *out_U, double *out_X,
                                   nsb200_stream_t stream) {
    if (check_model(model, true)) return 1;
    if (!sample_key || !out_U) return fail("NULL pointer");
    if (begin < 0 || end > N || begin > end) return fail("bad range [%lld, %lld) of %lld", (long long) begin, (long long) end, (long long) N);
    if (round < 0) return fail("round must be >= 0");
    if (end == begin) return 0;
    Geometry g;
    if (pick_geometry(model->D, g)) return 1;
    InitProposeArgs a{prior_only(*model), Key{sample_key[0], sample_key[1]}, begin, end, round, need, out_U, out_X};
    const size_t smem = 8 * model_smem_doubles(a.model.family, a.model.D, g.G, g.DPL, 0);
    const int per_block = kThreadsPerBlock / g.G;
    NSB_DISPATCH_GEOM(g, {
        if (set_smem(k_init_propose<kG, kDPL>, smem)) return 1;
        k_init_propose<kG, kDPL><<<grid_for(end - begin, per_block), kThreadsPerBlock, smem, (cudaStream_t) stream>>>(a);
    });
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_transform_batch(const NsModelDesc *model, const double *U, int64_t n, double *out_X,
                                      nsb200_stream_t stream) {
    if (check_model(model, true)) return 1;
    if (n <= 0) return 0;
    if (!U || !out_X) return fail("NULL pointer");
    Geometry g;
    if (pick_geometry(model->D, g)) return 1;
    const NsModelDesc pm = prior_only(*model);
    const size_t smem = 8 * model_smem_doubles(pm.family, pm.D, g.G, g.DPL, 0);
    const int per_block = kThreadsPerBlock / g.G;
    NSB_DISPATCH_GEOM(g, {
        if (set_smem(k_transform<kG, kDPL>, smem)) return 1;
        k_transform<kG, kDPL><<<grid_for(n, per_block), kThreadsPerBlock, smem, (cudaStream_t) stream>>>(pm, U, n, out_X);
    });
    NSB_LAUNCH_CHECK();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// statistics
// -------------------------------------------------------------------------------------------------
static size_t tree_workspace_bytes(long long M) {
    const long long n = M + 1;
    const long long tiles = (n + kScanTile - 1) / kScanTile;
    return sort_workspace_bytes(n) + align256((size_t) n * 4) + align256((size_t) tiles * 4) + 512;
}

extern "C" int64_t nsb200_workspace_bytes(int32_t op, int64_t n) {
    if (n < 0) n = 0;
    switch (op) {
        case NSB200_WS_ARGSORT: return (int64_t) sort_workspace_bytes(n > 0 ? n : 1);
        case NSB200_WS_COUNT_CROSSED_EDGES: return (int64_t) tree_workspace_bytes(n);
        case NSB200_WS_EVIDENCE_STATS: return (int64_t) (3 * ev_gpart_stride(kEvMaxGrid) * 8 + 512);
        case NSB200_WS_LOGSUMEXP: return 256;
        default: return -1;
    }
}

extern "C" int nsb200_argsort_f64(const double *keys, int64_t n, int64_t *out_idx, void *workspace,
                                  int64_t workspace_bytes, nsb200_stream_t stream) {
    if (n <= 0) return 0;
    if (n >= (1ll << 32)) return fail("argsort supports n < 2^32");
    if (!keys || !out_idx || !workspace) return fail("NULL pointer");
    if (workspace_bytes < (int64_t) sort_workspace_bytes(n)) return fail("workspace too small: %lld < %zu", (long long) workspace_bytes, sort_workspace_bytes(n));
    cudaStream_t st = (cudaStream_t) stream;
    SortWorkspace w = carve_sort_workspace(workspace, n);
    k_sort_init<<<1, 32, 0, st>>>(w.ctl);
    k_sort_prep<<<grid_for(n, 256), 256, 0, st>>>(keys, n, 0, w.keys[0], w.vals[0], w.ctl);
    NSB_CUDA(radix_sort_pairs(w, n, st));
    k_vals_to_i64<<<grid_for(n, 256), 256, 0, st>>>(w.vals[0], w.vals[1], w.ctl, n, (long long *) out_idx);
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_count_crossed_edges(const int64_t *sender_node_idx, const double *log_L, int64_t M,
                                          int64_t num_samples, int64_t *out_samples_indices,
                                          int32_t *out_num_live_points, void *workspace, int64_t workspace_bytes,
                                          nsb200_stream_t stream) {
    if (M <= 0) return 0;
    if (M + 1 >= (1ll << 31)) return fail("count_crossed_edges supports M < 2^31 - 1");
    if (num_samples > M) return fail("num_samples (%lld) > M (%lld)", (long long) num_samples, (long long) M);
    if (!sender_node_idx || !log_L || !out_samples_indices || !out_num_live_points || !workspace) return fail("NULL pointer");
    if (workspace_bytes < (int64_t) tree_workspace_bytes(M)) return fail("workspace too small");
    cudaStream_t st = (cudaStream_t) stream;
    const long long n = M + 1;
    SortWorkspace w = carve_sort_workspace(workspace, n);
    char *p = (char *) w.ctl + align256(sizeof(SortCtl));
    int *outdeg = (int *) p;
    p += align256((size_t) n * 4);
    int *tile_sums = (int *) p;
    const int tiles = (int) ((n + kScanTile - 1) / kScanTile);
    k_sort_init<<<1, 32, 0, st>>>(w.ctl);
    k_sort_prep<<<grid_for(n, 256), 256, 0, st>>>(log_L, M, 1, w.keys[0], w.vals[0], w.ctl);
    NSB_CUDA(radix_sort_pairs(w, n, st));
    NSB_CUDA(cudaMemsetAsync(outdeg, 0, (size_t) n * 4, st));
    k_out_degree<<<grid_for(M, 256), 256, 0, st>>>((const long long *) sender_node_idx, M, outdeg);
    k_tree_tile_sums<<<tiles, kScanThreads, 0, st>>>(w.vals[0], w.vals[1], w.ctl, outdeg, n, tile_sums);
    k_scan_u32_excl<<<1, 1024, 0, st>>>((uint32_t *) tile_sums, tiles);
    k_tree_apply<<<tiles, kScanThreads, 0, st>>>(w.vals[0], w.vals[1], w.ctl, outdeg, n, tile_sums, M, num_samples,
                                                 (long long *) out_samples_indices, out_num_live_points);
    NSB_LAUNCH_CHECK();
    return 0;
}

static NsEvidenceCalc init_evidence_calc() {
    NsEvidenceCalc c;
    c.log_L = -INFINITY;
    c.log_X_mean = 0.0;
    c.log_X2_mean = 0.0;
    c.log_Z_mean = -INFINITY;
    c.log_ZX_mean = -INFINITY;
    c.log_Z2_mean = -INFINITY;
    c.log_dZ_mean = -INFINITY;
    c.log_dZ2_mean = -INFINITY;
    return c;
}

extern "C" int nsb200_evidence_stats(const NsEvidenceCalc *init, const double *log_L, const double *num_live_points,
                                     int64_t M, NsEvidenceCalc *out_final, double *out_per_sample, void *workspace,
                                     int64_t workspace_bytes, nsb200_stream_t stream) {
    if (M < 0) return fail("M < 0");
    if (M > 0 && (!log_L || !num_live_points)) return fail("NULL input");
    if (!out_final) return fail("out_final is NULL");
    if (!workspace || workspace_bytes < (int64_t) (3 * ev_gpart_stride(kEvMaxGrid) * 8 + 512))
        return fail("workspace too small (nsb200_workspace_bytes(NSB200_WS_EVIDENCE_STATS, M))");
    EvSeq q;
    q.la = log_L;
    q.na = num_live_points;
    q.len_a = M;
    q.n_const_a = 0.0;
    q.lb = nullptr;
    q.len_b = 0;
    q.n_start_b = 0.0;
    q.tabT = q.tabT2 = q.tabt = nullptr;
    q.tab_n = 0;
    EvOut o;
    o.mid = nullptr;
    o.mark = -1;
    o.fin = out_final;
    o.per_sample = out_per_sample;
    double *gpart = (double *) align256((size_t) workspace);
    cudaStream_t st = (cudaStream_t) stream;
    NsEvidenceCalc init_v = init ? *init : init_evidence_calc();
    // a thread should own at least ~8 elements before more CTAs pay for their barrier trips
    long long want = (M + (long long) kEvThreads * 8 - 1) / ((long long) kEvThreads * 8);
    static int max_coresident = 0;
    if (!max_coresident) {
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_evidence_stats_grid, kEvThreads, 0) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        max_coresident = sms;  // one CTA per SM: the scans are latency / FP64 bound, more resident CTAs only add barrier cost
        if (max_coresident > kEvMaxGrid) max_coresident = kEvMaxGrid;
    }
    if (want <= kEvCluster) {
        k_evidence_stats<<<kEvCluster, kEvThreads, 0, st>>>(q, init_v, o, gpart);
        NSB_LAUNCH_CHECK();
        return 0;
    }
    int ctas = want > max_coresident ? max_coresident : (int) want;
    unsigned *bar = (unsigned *) ((char *) gpart + 3 * ev_gpart_stride(kEvMaxGrid) * 8);
    NSB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
    void *args[] = {&q, &init_v, &o, &gpart, &bar};
    NSB_CUDA(cudaLaunchCooperativeKernel((void *) k_evidence_stats_grid, dim3((unsigned) ctas), dim3(kEvThreads), args, 0, st));
    return 0;
}

extern "C" int nsb200_logsumexp(const double *x, int64_t n, double *out, void *workspace, int64_t workspace_bytes,
                                nsb200_stream_t stream) {
    (void) workspace;
    (void) workspace_bytes;
    if (!out) return fail("out is NULL");
    if (n > 0 && !x) return fail("x is NULL");
    k_logsumexp<<<1, 1024, 0, (cudaStream_t) stream>>>(x, n, out);
    NSB_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsb200_sample_evidence(const uint32_t key[2], const double *num_live_points, const double *log_L,
                                     int64_t M, int64_t S, double *out, nsb200_stream_t stream) {
    if (!key || !out) return fail("NULL argument");
    if (M < 0 || S < 0) return fail("negative size");
    if (M > 0 && (!num_live_points || !log_L)) return fail("input pointer is NULL");
    if (S == 0) return 0;
    k_sample_evidence<<<(unsigned) S, 1024, 0, (cudaStream_t) stream>>>(Key{key[0], key[1]}, num_live_points, log_L, M, out);
    NSB_LAUNCH_CHECK();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// engine
// -------------------------------------------------------------------------------------------------
__global__ void k_fill_f64(double *p, long long n, double v) {
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) p[i] = v;
}

__global__ void k_init_ctl(DevCtl *ctl, Key key) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->key = key;
    ctl->body = 0;
    ctl->stream_key[0] = split_child(split_child(key, 0), 1);  // sample_key of body 0
    {
        const Key k1 = split_child(split_child(split_child(key, 0), 0), 0);  // state key of body 1
        ctl->stream_key[1] = split_child(split_child(k1, 0), 1);
    }
    ctl->stream_key[2] = Key{0, 0};
    ctl->next_idx = 0;
    ctl->num_samples = 0;
    ctl->iteration = 0;
    ctl->sample_key = Key{0, 0};
    ctl->contour = -__longlong_as_double(0x7FF0000000000000ll);
    ctl->disc_start = 0;
    ctl->ph_start = 0;
    ctl->sender = 0;
    ctl->active = 1;
    ctl->err = 0;
    ctl->done_iter = -1;
    ctl->spec_ran = 0;
    ctl->spec_disc_start = ctl->spec_ph_start = -1;
    ctl->job[0].armed = ctl->job[1].armed = 0;
    ctl->cur = 1;  // the init scatter writes live0 ("other" buffer of cur = 1)
}

__global__ void k_set_cur(DevCtl *ctl, int cur) {
    if (threadIdx.x == 0 && blockIdx.x == 0) ctl->cur = cur;
}

// packs k_draw outputs into packed rows [U[D], logL, nevals]
__global__ void k_pack_rows(const double *U, const double *logL, const long long *nev, long long n, int D,
                            double *packed, long long row_doubles) {
    const long long total = n * (D + 2);
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long) gridDim.x * blockDim.x) {
        const long long i = t / (D + 2);
        const int j = (int) (t - i * (D + 2));
        double v;
        if (j < D) v = U[i * D + j];
        else if (j == D) v = logL[i];
        else v = __longlong_as_double(nev[i]);
        packed[i * row_doubles + j] = v;
    }
}

struct NsEngine {
    NsEngineConfig cfg;
    int D = 0;
    long long N = 0, m = 0, k = 0, cap = 0;
    long long rows_per_rank = 0, row_doubles = 0, packed_rows = 0;
    LiveSet live[2];
    DeadStore dead;
    DevCtl *ctl = nullptr;
    NsRegister *reg = nullptr;       // device
    NsRegister *reg_host = nullptr;  // pinned
    DevCtl *ctl_host = nullptr;      // pinned
    volatile long long *progress = nullptr;  // pinned + mapped: [0] bodies completed, [1] done
    long long *progress_dev = nullptr;
    double *seed_table = nullptr;
    double *packed = nullptr;
    unsigned *rank = nullptr;
    // fused all-gather over CUDA IPC peer mappings (world_size > 1, after nsb200_engine_p2p_connect)
    bool p2p = false;
    double *packed_home = nullptr;               // the engine's own gather / init-packing buffer (`packed` outside p2p bodies)
    double *packed_buf[2] = {nullptr, nullptr};  // p2p gather buffers by body parity (peers store into these)
    unsigned long long *p2p_flags = nullptr;     // [8] arrival epochs written by the peers
    int *p2p_err = nullptr;
    size_t off_packed[2] = {0, 0}, off_flags = 0;  // byte offsets inside the arena (exported to the peers)
    void *peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *peer_packed[2][8] = {};
    unsigned long long *peer_flags[8] = {};
    unsigned long long *p2p_epoch_dev = nullptr;  // barriers passed (device counter, advanced by active bodies only)
    uint64_t *sorted_new = nullptr;  // new keys sorted in tiles of 1024 (large shells: k_merge_sort_tiles)
    unsigned *new_pos = nullptr;     // row index of each sorted entry
    // chain streams of this rank's chains (k_chain_streams), triple buffered: the streams of body i+2
    // are generated on `side` in the gap between the slice kernels of bodies i and i+1
    double *pre_dirs[3] = {nullptr, nullptr, nullptr};
    double *pre_us[3] = {nullptr, nullptr, nullptr};
    uint2 *pre_rkeys[3] = {nullptr, nullptr, nullptr};
    double *tabT = nullptr, *tabT2 = nullptr, *tabt = nullptr;  // n-dependent evidence terms, n <= N
    EpiScratch *epi = nullptr;
    double *alpha_tab = nullptr;
    cudaStream_t side = nullptr;
    // register update of body b on its own stream, next to the slice kernel of body b + 1 (DevCtl::done_iter)
    cudaStream_t epi_stream = nullptr;
    cudaEvent_t ev_adv = nullptr, ev_epi[2] = {nullptr, nullptr};
    int slot = 0;  // parity of the body between step_begin and step_end
    int epi_ctas = 8;
    bool no_spec = true;  // this run's bodies are sequential: the register update stays on the caller's stream
    cudaEvent_t ev_keys = nullptr, ev_streams[3] = {nullptr, nullptr, nullptr};
    long long body = 0;       // host mirror of the next body index (its streams live in buffer body % 3)
    NsTermCond tc;
    std::vector<void *> allocs;
    // profiling
    std::vector<std::pair<void **, size_t>> arena_slots;  // pointer location, byte offset in the arena
    size_t arena_bytes = 0;
    int arena_device = -1;
    bool arena_exported = false;  // its IPC handle went to peer processes: never freed (see ArenaEntry)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    double slice_ms = 0.0;
    long long slice_launches = 0, all_launches = 0;
    bool initialised = false;
    // Where the generator of body b + 2 runs relative to the slice kernel of body b (see enqueue_streams):
    //   0  side stream, released together with the slice kernel (placement is a race between the two grids)
    //   1  same stream, slice kernel launched as its programmatic dependent: disjoint SM partitions
    //   2  same stream, before the slice kernel, whole GPU each (serial)
    //   3  side stream, released when the slice kernel has finished: overlaps the merge / register kernels
    int gen_mode = 3;
    bool pdl = false;  // gen_mode == 1
    cudaEvent_t ev_slice = nullptr;
    bool tables_ready = false;  // seed table / evidence-term tables / alpha table depend only on (N, S): built once
    // family EXTERNAL: chain state of the split slice step (ns_split.cuh) for this rank's chains
    bool external = false;
    int grad_flags = 0;  // gradient_slice (1) / gradient_guided (2) chains: nsb200_engine_set_split_flags
    int split_P = 1;     // proposals per chain and round of the split step
    void *split_ws = nullptr;
    size_t split_ws_bytes = 0;
};

// Engine buffers are carved out of ONE device allocation: dev_alloc only records (where the pointer lives, offset),
// arena_commit makes the single cudaMalloc and fills the pointers in.  (30 separate cudaMalloc calls cost ~5 ms per
// NestedSampler, 5 % of a config-2 run through the public API.)
template <typename T>
static int dev_alloc(NsEngine *e, T **p, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t) 255;
    e->arena_slots.push_back({(void **) p, e->arena_bytes});
    e->arena_bytes += bytes;
    *p = nullptr;
    return 0;
}

// Arenas of destroyed engines are kept in a small per-process pool and handed to the next engine of exactly the
// same size: a sampler rebuilt for every run -- what the public API invites -- then pays no cudaMalloc / cudaFree
// (each 1-10 ms of driver time for a 100 MB arena, and a device-wide synchronisation).  An arena whose CUDA IPC
// handle has been given to peer processes (fused all-gather) is never freed: the peers keep their mappings for the
// life of the process (see open_peer_arena), and a rebuilt engine of the same size gets the same arena, so its
// handle -- and the peers' mappings -- stay valid without any re-wiring.
struct ArenaEntry {
    void *ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
    bool exported = false;
};
static std::vector<ArenaEntry> g_arena_pool;
static std::mutex g_arena_mutex;
constexpr size_t kArenaPoolMax = 4;

static int arena_commit(NsEngine *e) {
    void *q = nullptr;
    int dev = -1;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_arena_mutex);
        for (size_t i = 0; i < g_arena_pool.size(); ++i) {
            if (g_arena_pool[i].bytes == e->arena_bytes && g_arena_pool[i].device == dev) {
                q = g_arena_pool[i].ptr;
                e->arena_exported = g_arena_pool[i].exported;
                g_arena_pool.erase(g_arena_pool.begin() + i);
                break;
            }
        }
    }
    if (!q) {
        cudaError_t err = cudaMalloc(&q, e->arena_bytes);
        if (err != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", e->arena_bytes, cudaGetErrorString(err));
    }
    e->arena_device = dev;
    e->allocs.push_back(q);
    for (const auto &sl : e->arena_slots) *sl.first = (char *) q + sl.second;
    return 0;
}

// Peer arenas opened through CUDA IPC, by handle: opened once per process and never closed (opening costs ~15 ms
// and closing the last mapping of a peer device tears its peer access down -- 100 ms stalls in the middle of a run).
struct PeerMapping {
    uint8_t handle[64];
    void *ptr;
};
static std::vector<PeerMapping> g_peer_maps;
static std::mutex g_peer_mutex;

static int open_peer_arena(const uint8_t *handle, void **out) {
    std::lock_guard<std::mutex> lock(g_peer_mutex);
    for (const auto &m : g_peer_maps) {
        if (memcmp(m.handle, handle, 64) == 0) {
            *out = m.ptr;
            return 0;
        }
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *q = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(err));
    }
    PeerMapping m;
    memcpy(m.handle, handle, 64);
    m.ptr = q;
    g_peer_maps.push_back(m);
    *out = q;
    return 0;
}

extern "C" void nsb200_engine_destroy(NsEngine *e) {
    if (!e) return;
    for (void *p : e->allocs) {
        // kernels of this engine may still be running on the caller's stream: the next owner of the arena is only
        // safe after they are done (cudaFree would have waited for them as well)
        cudaDeviceSynchronize();
        void *evict = p;
        {
            std::lock_guard<std::mutex> lock(g_arena_mutex);
            if (e->arena_bytes && e->arena_device >= 0) {
                evict = nullptr;
                ArenaEntry a;
                a.ptr = p;
                a.bytes = e->arena_bytes;
                a.device = e->arena_device;
                a.exported = e->arena_exported;
                g_arena_pool.push_back(a);
                if (g_arena_pool.size() > kArenaPoolMax) {  // evict the oldest arena no peer can be mapping
                    for (size_t i = 0; i < g_arena_pool.size(); ++i) {
                        if (!g_arena_pool[i].exported) {
                            evict = g_arena_pool[i].ptr;
                            g_arena_pool.erase(g_arena_pool.begin() + i);
                            break;
                        }
                    }
                }
            }
        }
        if (evict) cudaFree(evict);
    }
    if (e->reg_host) cudaFreeHost(e->reg_host);
    if (e->ctl_host) cudaFreeHost(e->ctl_host);
    if (e->progress) cudaFreeHost((void *) e->progress);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->ev_keys) cudaEventDestroy(e->ev_keys);
    if (e->ev_slice) cudaEventDestroy(e->ev_slice);
    for (int b = 0; b < 3; ++b)
        if (e->ev_streams[b]) cudaEventDestroy(e->ev_streams[b]);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->ev_adv) cudaEventDestroy(e->ev_adv);
    for (int b = 0; b < 2; ++b)
        if (e->ev_epi[b]) cudaEventDestroy(e->ev_epi[b]);
    if (e->epi_stream) cudaStreamDestroy(e->epi_stream);
    delete e;
}

extern "C" int nsb200_engine_create(const NsEngineConfig *cfg, NsEngine **out) {
    if (!cfg || !out) return fail("NULL argument");
    if (check_model(&cfg->model, true)) return 1;
    Geometry g;
    if (pick_geometry(cfg->model.D, g)) return 1;
    if (cfg->num_live_points < 2) return fail("num_live_points must be >= 2");
    if (cfg->shell_size < 1 || cfg->shell_size > cfg->num_live_points) return fail("bad shell_size %lld", (long long) cfg->shell_size);
    if (cfg->num_slices < 1) return fail("num_slices should be >= 1, got %d", cfg->num_slices);
    if (cfg->num_phantom < 0 || cfg->num_phantom >= cfg->num_slices)
        return fail("expected 0 <= num_phantom < num_slices, got %d, %d", cfg->num_phantom, cfg->num_slices);
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return fail("bad rank/world_size");
    if (cfg->shell_size % cfg->world_size != 0)
        return fail("shell_size %lld is not a multiple of world_size %d (round_up_num_live_points)", (long long) cfg->shell_size, cfg->world_size);
    const long long block = cfg->shell_size * (1 + (long long) cfg->num_phantom);
    if (cfg->max_samples < cfg->num_live_points || cfg->max_samples < block)
        return fail("max_samples %lld too small", (long long) cfg->max_samples);
    if (cfg->num_live_points >= (1ll << 31)) return fail("num_live_points too large");
    NsEngine *e = new NsEngine();
    e->cfg = *cfg;
    e->D = cfg->model.D;
    e->N = cfg->num_live_points;
    e->m = cfg->shell_size;
    e->k = cfg->num_phantom;
    e->cap = cfg->max_samples;
    e->rows_per_rank = e->m / cfg->world_size;
    e->row_doubles = (e->D + 2) + e->k * (e->D + 1);
    e->packed_rows = e->N > e->m ? e->N : e->m;  // the init pass packs all N prior draws
    e->external = cfg->model.family == NSB200_FAM_EXTERNAL;
    e->gen_mode = opt(OPT_GEN_MODE);
    if (e->gen_mode < 0 || e->gen_mode > 4) e->gen_mode = 3;
    e->pdl = e->gen_mode == 1;
    const size_t D = e->D;
    int rc = 0;
    for (int b = 0; b < 2 && !rc; ++b) {
        rc |= dev_alloc(e, &e->live[b].sender, e->N);
        rc |= dev_alloc(e, &e->live[b].U, e->N * D);
        rc |= dev_alloc(e, &e->live[b].logL_constraint, e->N);
        rc |= dev_alloc(e, &e->live[b].logL, e->N);
        rc |= dev_alloc(e, &e->live[b].nevals, e->N);
    }
    e->dead.capacity = e->cap;
    if (!rc) rc |= dev_alloc(e, &e->dead.sender, e->cap);
    if (!rc) rc |= dev_alloc(e, &e->dead.logL, e->cap);
    if (!rc) rc |= dev_alloc(e, &e->dead.U, e->cap * D);
    if (!rc) rc |= dev_alloc(e, &e->dead.nevals, e->cap);
    if (!rc) rc |= dev_alloc(e, &e->dead.phantom, e->cap);
    if (!rc) rc |= dev_alloc(e, &e->ctl, 1);
    if (!rc) rc |= dev_alloc(e, &e->reg, 1);
    if (!rc) rc |= dev_alloc(e, &e->seed_table, e->N);
    if (!rc) rc |= dev_alloc(e, &e->packed, (size_t) e->packed_rows * e->row_doubles);
    if (cfg->world_size > 1 && cfg->world_size <= 8) {
        for (int b = 0; b < 2; ++b) {
            e->off_packed[b] = e->arena_bytes;
            if (!rc) rc |= dev_alloc(e, &e->packed_buf[b], (size_t) e->m * e->row_doubles);
        }
        e->off_flags = e->arena_bytes;
        if (!rc) rc |= dev_alloc(e, &e->p2p_flags, 16);  // [0..8) arrival epochs, [8..16) the peers' contours
        if (!rc) rc |= dev_alloc(e, &e->p2p_epoch_dev, 1);
        if (!rc) rc |= dev_alloc(e, &e->p2p_err, 1);
    }
    if (!rc) rc |= dev_alloc(e, &e->rank, e->N);
    if (!rc) rc |= dev_alloc(e, &e->sorted_new, e->N + kMergeTile);  // whole tiles (pads included)
    if (!rc) rc |= dev_alloc(e, &e->new_pos, e->N + kMergeTile);
    if (!rc) rc |= dev_alloc(e, &e->epi, 1);
    if (!rc) rc |= dev_alloc(e, &e->alpha_tab, cfg->num_slices);
    if (!rc) rc |= dev_alloc(e, &e->tabT, e->N + 2);
    if (!rc) rc |= dev_alloc(e, &e->tabT2, e->N + 2);
    if (!rc) rc |= dev_alloc(e, &e->tabt, e->N + 2);
    if (e->external) {
        e->split_ws_bytes = split_workspace_bytes(e->D, e->rows_per_rank, (int) e->k);
        if (!rc) rc |= dev_alloc(e, (char **) &e->split_ws, e->split_ws_bytes);
    } else if (g.G >= 8) {  // data-independent chain streams are generated off the chains' critical path
        const size_t rows = (size_t) e->rows_per_rank * cfg->num_slices;
        for (int b = 0; b < 3; ++b) {
            if (!rc) rc |= dev_alloc(e, &e->pre_dirs[b], rows * D);
            if (!rc) rc |= dev_alloc(e, &e->pre_us[b], rows * kPre);
            if (!rc) rc |= dev_alloc(e, &e->pre_rkeys[b], rows);
        }
        if (!rc && cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) != cudaSuccess) rc = fail("cudaStreamCreate failed");
        if (!rc && cudaEventCreateWithFlags(&e->ev_keys, cudaEventDisableTiming) != cudaSuccess) rc = fail("cudaEventCreate failed");
        if (!rc && cudaEventCreateWithFlags(&e->ev_slice, cudaEventDisableTiming) != cudaSuccess) rc = fail("cudaEventCreate failed");
        for (int b = 0; b < 3; ++b)
            if (!rc && cudaEventCreateWithFlags(&e->ev_streams[b], cudaEventDisableTiming) != cudaSuccess) rc = fail("cudaEventCreate failed");
    }
    {
        // the register update has 8-64 CTAs that meet at software barriers: highest priority, so that they are placed
        // ahead of the thousands of chain CTAs of the next body that are enqueued right behind them
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (!rc && cudaStreamCreateWithPriority(&e->epi_stream, cudaStreamNonBlocking, opt(OPT_EPI_PRIO) ? prio_hi : prio_lo) != cudaSuccess)
            rc = fail("cudaStreamCreate failed");
    }
    // CTAs of the in-loop register update: ~6 elements of the m + N scanned per thread (the 8-element register cache
    // of evidence_scan_block), between the 8 of config 2 and 64
    {
        // ... but never more than the SMs the chain-stream generator leaves free (enqueue_streams: 28 while it is
        // short, 8 when it is long): the CTAs meet at a spinning barrier, so all of them have to be resident at once --
        // 49 CTAs at config 5 sat behind the 5 ms generator and cost 17 ms per body (profiles/r2/config5_n8_49cta_epilogue.json)
        const long long want = (e->m + e->N + (long long) kEvThreads * 6 - 1) / ((long long) kEvThreads * 6);
        const bool short_gen = (double) e->rows_per_rank * cfg->num_slices * e->D < 32e6;
        const long long room = short_gen ? 24 : kEvCluster;
        e->epi_ctas = (int) (want < kEvCluster ? kEvCluster : (want > room ? room : want));
        if (opt(OPT_EPI_CTAS) >= 1 && opt(OPT_EPI_CTAS) <= 64) e->epi_ctas = opt(OPT_EPI_CTAS);
    }
    if (!rc && cudaEventCreateWithFlags(&e->ev_adv, cudaEventDisableTiming) != cudaSuccess) rc = fail("cudaEventCreate failed");
    for (int b = 0; b < 2; ++b)
        if (!rc && cudaEventCreateWithFlags(&e->ev_epi[b], cudaEventDisableTiming) != cudaSuccess) rc = fail("cudaEventCreate failed");
    if (!rc) rc |= arena_commit(e);
    if (!rc) e->packed_home = e->packed;
    if (!rc && e->p2p_flags) {
        if (cudaMemset(e->p2p_flags, 0, 16 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMemset(e->p2p_epoch_dev, 0, sizeof(unsigned long long)) != cudaSuccess ||
            cudaMemset(e->p2p_err, 0, sizeof(int)) != cudaSuccess)
            rc = fail("cudaMemset failed");
    }
    if (!rc && cudaMallocHost((void **) &e->reg_host, sizeof(NsRegister)) != cudaSuccess) rc = fail("cudaMallocHost failed");
    if (!rc && cudaMallocHost((void **) &e->ctl_host, sizeof(DevCtl)) != cudaSuccess) rc = fail("cudaMallocHost failed");
    if (!rc) {
        void *hp = nullptr;
        if (cudaHostAlloc(&hp, 64, cudaHostAllocMapped) != cudaSuccess) rc = fail("cudaHostAlloc(mapped) failed");
        else {
            e->progress = (volatile long long *) hp;
            e->progress[0] = -1;
            e->progress[1] = 0;
            e->progress[2] = 0;
            if (cudaHostGetDevicePointer((void **) &e->progress_dev, hp, 0) != cudaSuccess) rc = fail("cudaHostGetDevicePointer failed");
        }
    }
    if (rc) {
        nsb200_engine_destroy(e);
        return 1;
    }
    *out = e;
    return 0;
}

// ---- fused all-gather: CUDA IPC wiring (host side exchanges the handles, e.g. with all_gather_object) ---------
extern "C" int nsb200_engine_p2p_export(NsEngine *e, uint8_t handle[64], int64_t offsets[3]) {
    if (!e || !handle || !offsets) return fail("NULL argument");
    if (!e->p2p_flags) return fail("p2p needs 2 <= world_size <= 8");
    if (e->external) return fail("family EXTERNAL uses the host all-gather");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    NSB_CUDA(cudaIpcGetMemHandle(&h, e->allocs[0]));
    e->arena_exported = true;
    memcpy(handle, &h, 64);
    offsets[0] = (int64_t) e->off_packed[0];
    offsets[1] = (int64_t) e->off_packed[1];
    offsets[2] = (int64_t) e->off_flags;
    return 0;
}

extern "C" int nsb200_engine_p2p_connect(NsEngine *e, const uint8_t *handles, const int64_t *offsets) {
    if (!e || !handles || !offsets) return fail("NULL argument");
    if (!e->p2p_flags) return fail("p2p needs 2 <= world_size <= 8");
    const int world = e->cfg.world_size, me = e->cfg.rank;
    for (int r = 0; r < world; ++r) {
        char *base;
        if (r == me) {
            base = (char *) e->allocs[0];
        } else {
            void *q = nullptr;
            if (open_peer_arena(handles + (size_t) r * 64, &q)) return 1;
            e->peer_base[r] = q;
            base = (char *) q;
        }
        e->peer_packed[0][r] = (double *) (base + offsets[r * 3 + 0]);
        e->peer_packed[1][r] = (double *) (base + offsets[r * 3 + 1]);
        e->peer_flags[r] = (unsigned long long *) (base + offsets[r * 3 + 2]);
    }
    e->p2p = true;
    return 0;
}

extern "C" int nsb200_engine_p2p_enabled(NsEngine *e, int32_t enable) {
    if (!e) return fail("NULL engine");
    if (enable >= 0) {
        if (enable && !e->peer_flags[e->cfg.rank]) return fail("p2p is not connected");
        e->p2p = enable != 0;
    }
    return 0;
}

extern "C" int nsb200_engine_p2p_error(NsEngine *e, int32_t *out) {
    if (!e || !out) return fail("NULL argument");
    *out = 0;
    if (e->p2p_err) {
        int v = 0;
        NSB_CUDA(cudaMemcpy(&v, e->p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
        *out = v;
    }
    return 0;
}

static NsRegister init_register_host() {
    NsRegister r;
    memset(&r, 0, sizeof(r));
    r.evidence_calc = init_evidence_calc();
    r.evidence_calc_with_remaining = init_evidence_calc();
    r.log_L_contour = -INFINITY;
    r.efficiency = 0.0;
    r.relative_spread = INFINITY;
    r.absolute_spread = INFINITY;
    r.peak_log_XL = -INFINITY;
    return r;
}

// _main_ns_thread lowers max_samples by one iteration's space (sharded_static.py:464-470).
static NsTermCond effective_term_cond(const NsEngine *e, const NsTermCond *tc) {
    NsTermCond t;
    if (tc) t = *tc;
    else {
        memset(&t, 0, sizeof(t));
        t.mask = (1u << 3) | (1u << 4);
        t.dlogZ = log(1.0 + 1e-3);
        t.max_samples = (double) e->cap;
    }
    if (t.mask & (1u << 4)) {
        const double lim = (double) (e->cap - e->m * (1 + e->k));
        if (t.max_samples > lim) t.max_samples = lim;
    }
    return t;
}

static int enqueue_streams(NsEngine *e, int buf, cudaStream_t st, cudaEvent_t after);

static void launch_merge_rank(NsEngine *e, long long m_new, cudaStream_t st) {
    // NSB200_MERGE_BRUTE: A/B and parity knob, results do not depend on it
    if (!opt(OPT_MERGE_BRUTE) && m_new >= 2048) {
        const unsigned tiles = (unsigned) ((m_new + kMergeTile - 1) / kMergeTile);
        k_merge_sort_tiles<<<tiles, kMergeTile, 0, st>>>(e->ctl, e->packed, e->row_doubles, e->D, m_new, e->sorted_new,
                                                         e->new_pos);
        k_merge_rank_tiles<<<grid_for(e->N * kMergeLanes, 256), 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->packed,
                                                                               e->row_doubles, e->D, m_new, e->N,
                                                                               e->sorted_new, e->new_pos, e->rank);
        e->all_launches += 1;
        return;
    }
    if (e->N <= 16384)
        k_merge_rank<32><<<grid_for(e->N * 32, 256), 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->packed, e->row_doubles,
                                                                  e->D, m_new, e->N, e->rank);
    else
        k_merge_rank<8><<<grid_for(e->N * 8, 256), 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->packed, e->row_doubles,
                                                                e->D, m_new, e->N, e->rank);
}

static int engine_init_common(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond, nsb200_stream_t stream,
                              const double *extU, const double *extL, const long long *extN) {
    if (!e || !key) return fail("NULL argument");
    cudaStream_t st = (cudaStream_t) stream;
    const int D = e->D;
    e->tc = effective_term_cond(e, term_cond);
    e->progress[0] = -1;
    e->progress[1] = 0;
    e->progress[2] = 0;
    e->slice_ms = 0.0;
    e->slice_launches = 0;
    e->all_launches = 0;
    e->ev_used = 0;
    e->packed = e->packed_home;
    e->body = 0;
    NSB_CUDA(cudaStreamSynchronize(e->epi_stream));  // register updates of a previous run on this engine
    if (e->p2p) {  // run-entry barrier: peers may only store rows of this run once every rank has left the previous one
        PeerFlags pf;
        for (int r = 0; r < 8; ++r) pf.p[r] = e->peer_flags[r];
        k_peer_barrier<<<1, 32, 0, st>>>(e->ctl, e->p2p_epoch_dev, e->p2p_flags, pf, e->cfg.world_size, e->cfg.rank, e->p2p_err, 1, e->progress_dev);
    }
    // create_init_state (initialisation.py:38-47): empty dead store, key split
    NSB_CUDA(cudaMemsetAsync(e->dead.sender, 0, (size_t) e->cap * 8, st));
    NSB_CUDA(cudaMemsetAsync(e->dead.U, 0, (size_t) e->cap * D * 8, st));
    NSB_CUDA(cudaMemsetAsync(e->dead.nevals, 0, (size_t) e->cap * 8, st));
    NSB_CUDA(cudaMemsetAsync(e->dead.phantom, 0, (size_t) e->cap, st));
    k_fill_f64<<<592, 256, 0, st>>>(e->dead.logL, e->cap, INFINITY);
    const Key k0{key[0], key[1]};
    const Key key1 = split_child(k0, 0), sample_key = split_child(k0, 1);
    k_init_ctl<<<1, 1, 0, st>>>(e->ctl, key1);
    if (e->pre_dirs[0]) {
        // the side stream may still be writing streams from a previous run into these buffers
        NSB_CUDA(cudaStreamSynchronize(e->side));
        NSB_CUDA(cudaEventRecord(e->ev_keys, st));
        if (enqueue_streams(e, 0, st, e->ev_keys)) return 1;  // bodies 0 and 1; body b + 2 is enqueued by body b
        if (enqueue_streams(e, 1, st, e->ev_keys)) return 1;
    }
    const bool build_tables = !e->tables_ready;
    if (build_tables) {
        // The seed table is a serial logaddexp recurrence (it has to round like the reference's sequential
        // cumulative_logsumexp: 36 us per 100 entries); it only depends on N, so repeated runs reuse it.
        if (nsb200_seed_table(e->N, e->seed_table, stream)) return 1;
        k_ev_tables<<<64, 256, 0, st>>>(e->N, e->tabT, e->tabT2, e->tabt);
        k_alpha_table<<<4, 256, 0, st>>>(e->cfg.num_slices, e->alpha_tab);
        e->tables_ready = true;
    }
    // N prior draws (replicated on every rank), packed, ranked (stable argsort) and scattered
    const double *tmpU = e->live[1].U, *tmpL = e->live[1].logL;
    const long long *tmpN = e->live[1].nevals;
    if (extU) {  // caller-evaluated initial live points (family EXTERNAL)
        tmpU = extU;
        tmpL = extL;
        tmpN = extN;
    } else if (launch_draw(&e->cfg.model, sample_key, nullptr, 0, e->N, e->live[1].U, e->live[1].logL, e->live[1].nevals, 0, st)) {
        return 1;
    }
    k_pack_rows<<<592, 256, 0, st>>>(tmpU, tmpL, tmpN, e->N, D, e->packed, e->row_doubles);
    launch_merge_rank(e, e->N, st);
    DeadStore nodead = e->dead;
    k_merge_scatter<<<592, 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->packed, e->row_doubles, D, e->N, e->N, 0,
                                          e->rank, nodead, 0);
    k_set_cur<<<1, 1, 0, st>>>(e->ctl, 0);
    // create_init_termination_register + the loop-entry no_seed_points and first cond
    *e->reg_host = init_register_host();
    NSB_CUDA(cudaMemcpyAsync(e->reg, e->reg_host, sizeof(NsRegister), cudaMemcpyHostToDevice, st));
    k_iter_epilogue<<<kEvCluster, kEvThreads, 0, st>>>(e->ctl, e->reg, e->live[0], e->live[1], e->packed, e->row_doubles, D, e->m,
                                         e->N, e->tc, 1, e->tabT, e->tabT2, e->tabt, e->N, e->epi, e->progress_dev);
    NSB_LAUNCH_CHECK();
    e->all_launches += build_tables ? 9 : 6;
    e->initialised = true;
    return 0;
}

extern "C" int nsb200_engine_init(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond,
                                  nsb200_stream_t stream) {
    if (e && e->external) return fail("family EXTERNAL: use nsb200_engine_init_external with caller-evaluated live points");
    return engine_init_common(e, key, term_cond, stream, nullptr, nullptr, nullptr);
}

extern "C" int nsb200_engine_init_external(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond,
                                           const double *U, const double *log_L,
                                           const int64_t *num_likelihood_evaluations, nsb200_stream_t stream) {
    if (!U || !log_L || !num_likelihood_evaluations) return fail("initial live points are NULL");
    return engine_init_common(e, key, term_cond, stream, U, log_L, (const long long *) num_likelihood_evaluations);
}

// NSB200_TRACE=1: device timeline of a few iterations (debug aid for DESIGN.md numbers)
struct TraceMark { cudaEvent_t ev; const char *name; };
static std::vector<TraceMark> g_trace;
static bool trace_on(NsEngine *e) {
    return opt(OPT_TRACE) && e->slice_launches >= 60 && e->slice_launches < 64;
}
static void trace_mark(NsEngine *e, const char *name, cudaStream_t s) {
    if (!trace_on(e)) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, s);
    g_trace.push_back({ev, name});
}
static void trace_dump() {
    if (g_trace.empty()) return;
    cudaDeviceSynchronize();
    for (size_t i = 1; i < g_trace.size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_trace[0].ev, g_trace[i].ev);
        fprintf(stderr, "trace %8.1f us  %s\n", ms * 1e3, g_trace[i].name);
    }
    g_trace.clear();
}

static cudaEvent_t next_event(NsEngine *e) {
    if (e->ev_used == e->ev_pool.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->ev_pool.push_back(ev);
    }
    return e->ev_pool[e->ev_used++];
}

// resolves pending event pairs into slice_ms (requires the stream to be idle)
static void drain_events(NsEngine *e) {
    for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e->ev_pool[i], e->ev_pool[i + 1]) == cudaSuccess) e->slice_ms += ms;
    }
    e->ev_used = 0;
}

// Enqueues the generation of the chain streams of the body whose sample_key is ctl->stream_key[buf] into
// buffer `buf`; `after` = event the side stream waits for first (gen_mode 0 and 3).  See NsEngine::gen_mode.
static int enqueue_streams(NsEngine *e, int buf, cudaStream_t st, cudaEvent_t after) {
    const bool own_stream = e->gen_mode == 0 || e->gen_mode >= 3;
    cudaStream_t gs = own_stream ? e->side : st;
    if (own_stream) NSB_CUDA(cudaStreamWaitEvent(e->side, after, 0));
    const long long begin = e->rows_per_rank * e->cfg.rank, end = begin + e->rows_per_rank;
    StreamArgs sa;
    sa.key = Key{0, 0};
    sa.ctl = e->ctl;
    sa.key_slot = buf;
    sa.chain_begin = begin;
    sa.chain_end = end;
    sa.S = e->cfg.num_slices;
    sa.D = e->D;
    sa.dirs = e->pre_dirs[buf];
    sa.us = e->pre_us[buf];
    sa.rkeys = e->pre_rkeys[buf];
    const long long warps = (end - begin) * ((sa.S + 31) / 32);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int gen_ctas, tpb;
    size_t gen_smem = 0;
    if (e->gen_mode <= 1) {
        // The generator is throughput-bound, the chains are latency-bound: sharing an SM starves the chains.
        // Partition modes give the generator its own SMs: persistent CTAs of 1024 threads that each claim ALL
        // shared memory of an SM, so that no slice CTA can land next to them.
        gen_ctas = (sms * 43) / 100;
        tpb = 1024;
        static size_t optin_smem = 0;
        if (!optin_smem) {
            int optin = 0;
            optin_smem = 227 * 1024;
            if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0) == cudaSuccess && optin > 0) optin_smem = (size_t) optin;
            NSB_CUDA(cudaFuncSetAttribute(k_chain_streams<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) optin_smem));
            NSB_CUDA(cudaFuncSetAttribute(k_chain_streams<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) optin_smem));
            NSB_CUDA(cudaFuncSetAttribute(k_chain_streams<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) optin_smem));
        }
        gen_smem = e->gen_mode == 1 ? optin_smem : 200 * 1024;
    } else {
        // Whole-GPU modes: one persistent CTA of 32 warps per SM (measured: 16 warps per SM lose more generator
        // throughput than the room they leave for the merge / register-update kernels gains).
        // mode 3: the merge / register-update kernels of step_end run next to the generator; the register update
        // is a cluster of 8 CTAs that needs whole free SMs.  While the generator is short (config 2: ~150 us, about
        // the length of step_end) leaving 28 SMs measured 3 % faster per run than leaving 8: it runs longer but off
        // the critical path.  A long generator (config 5: D = 100, S = 500) is exposed anyway and wants every SM
        // (profiles/r1/gen_sweep_r1.txt, config_sweep_r1.txt).
        const bool short_gen = (double) (end - begin) * sa.S * sa.D < 32e6;
        gen_ctas = e->gen_mode == 3 ? (short_gen && sms > 56 ? sms - 28 : sms - 8) : sms;
        tpb = 1024;
        if (e->gen_mode == 4) tpb = 128;  // back-fill: one small CTA per SM in the registers the chains leave free
    }
    if (opt(OPT_GEN_SMS) > 0) gen_ctas = opt(OPT_GEN_SMS);
    if (opt(OPT_GEN_TPB) >= 32 && opt(OPT_GEN_TPB) <= 1024) tpb = (opt(OPT_GEN_TPB) / 32) * 32;
    const long long wpc = tpb / 32;
    (void) wpc;
    if (launch_chain_streams(sa, gen_ctas, tpb, gen_smem, gs)) return 1;
    if (own_stream) {
        NSB_CUDA(cudaEventRecord(e->ev_streams[buf], e->side));
        trace_mark(e, "  generator end (side)", e->side);
    }
    e->all_launches += 1;
    return 0;
}

extern "C" int nsb200_engine_step_begin(NsEngine *e, nsb200_stream_t stream) {
    if (!e || !e->initialised) return fail("engine not initialised");
    cudaStream_t st = (cudaStream_t) stream;
    const int D = e->D;
    // streams of THIS body were enqueued two steps ago (or by init) on the side stream
    trace_mark(e, "step_begin enqueue", st);
    const bool own_stream = e->gen_mode == 0 || e->gen_mode >= 3;
    if (e->pre_dirs[0] && own_stream) NSB_CUDA(cudaStreamWaitEvent(st, e->ev_streams[e->body % 3], 0));
    trace_mark(e, "after wait streams", st);
    // the loop condition this body starts under is the register of the body before the previous one (same slot):
    // the previous body's register update may still be running next to this body's chains (DevCtl::done_iter)
    e->slot = (int) (e->body & 1);
    NSB_CUDA(cudaStreamWaitEvent(st, e->ev_epi[e->slot], 0));
    // A store that wraps around (no max_samples bound: SimpleGlobalOptimisation, sharded_static.py:76-78) cannot be
    // rolled back -- the speculative body would overwrite the oldest rows of the ring -- so such runs are sequential
    // NSB200_SPECULATE: 1 = always overlap the register update with the next body, 0 = never, 2 (default) = when the
    // live set is replicated over several GPUs.  On one GPU the update (75 us) already hides behind the chain-stream
    // generator (139 us on the side stream), so overlapping it with the next chains only makes chains and generator
    // collide: same 93 ms per config-2 run either way, and one wasted body at the end.  From 2 GPUs on the update
    // (100-140 us over the larger replicated live set) is the longer of the two and comes off the critical path:
    // 8-GPU weak scaling 0.79 -> 0.92 (profiles/r2/).
    const int spec_mode = opt(OPT_SPECULATE);
    const bool speculate = spec_mode == 1 || (spec_mode == 2 && e->cfg.world_size > 1);
    const int no_spec = ((e->tc.mask & (1u << 4)) && speculate) ? 0 : 1;
    e->no_spec = no_spec != 0;
    if (no_spec) NSB_CUDA(cudaStreamWaitEvent(st, e->ev_epi[e->slot ^ 1], 0));
    // NSB200_GEN_FENCE=1: the generator launched behind the previous slice kernel (streams of body + 1) gets the GPU to
    // itself before this body's chains start.  Measured slower (105 vs 93 ms per config-2 run): with the register
    // update off the main stream the chains of body b + 1 simply start under the generator's tail.
    if (e->pre_dirs[0] && e->gen_mode == 3 && opt(OPT_GEN_FENCE) && e->body >= 1)
        NSB_CUDA(cudaStreamWaitEvent(st, e->ev_streams[(e->body + 1) % 3], 0));
    k_iter_prologue<<<1, 1, 0, st>>>(e->ctl, e->reg, e->live[0], e->live[1], e->m, e->k, e->cap, e->cfg.intended_sender, e->epi,
                                     no_spec);
    k_append_live<<<296, 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->dead, e->m, D, 0);
    NSB_LAUNCH_CHECK();
    if (e->external) {  // the chains are run by the caller through nsb200_engine_split_*
        e->body += 1;
        e->all_launches += 2;
        return 0;
    }
    // this rank's chains -> its block of the gather buffer
    const long long begin = e->rows_per_rank * e->cfg.rank, end = begin + e->rows_per_rank;
    SliceArgs a;
    memset(&a, 0, sizeof(a));
    a.model = e->cfg.model;
    a.seed_table = e->seed_table;
    a.N = e->N;
    a.chain_begin = begin;
    a.chain_end = end;
    a.S = e->cfg.num_slices;
    a.k = e->cfg.num_phantom;
    a.midpoint = e->cfg.midpoint_shrink;
    a.packed = e->packed + begin * e->row_doubles;
    a.packed_row_doubles = e->row_doubles;
    if (e->p2p) {
        // gather buffers alternate with the body parity (bodies are numbered from 0 in every run, identically on
        // every rank): a peer one body ahead writes the other buffer
        const int par = (int) (e->body & 1);
        e->packed = e->packed_buf[par];
        a.packed = nullptr;
        a.n_peers = e->cfg.world_size;
        for (int r = 0; r < a.n_peers; ++r) a.peers[r] = e->peer_packed[par][r] + begin * e->row_doubles;
    }
    a.ctl = e->ctl;
    a.err = &e->ctl->err;
    a.live0 = e->live[0];
    a.live1 = e->live[1];
    a.alpha_tab = e->alpha_tab;
    const int buf = (int) (e->body % 3);
    a.pre_dirs = e->pre_dirs[buf];
    a.pre_us = e->pre_us[buf];
    a.pre_rkeys = e->pre_rkeys[buf];
    cudaEvent_t e0 = next_event(e), e1 = next_event(e);
    const int nbuf = (int) ((e->body + 2) % 3);  // streams of body + 2
    if (e->pdl) {  // nothing may sit between the generator and its programmatic dependent in the stream
        cudaEventRecord(e0, st);
        trace_mark(e, "slice start", st);
    }
    if (e->pre_dirs[0] && e->gen_mode == 0) {
        NSB_CUDA(cudaEventRecord(e->ev_keys, st));
        if (enqueue_streams(e, nbuf, st, e->ev_keys)) return 1;
    }
    if (e->pre_dirs[0] && (e->gen_mode == 1 || e->gen_mode == 2) && enqueue_streams(e, nbuf, st, nullptr)) return 1;
    if (e->pre_dirs[0] && e->gen_mode == 4) NSB_CUDA(cudaEventRecord(e->ev_keys, st));
    if (!e->pdl) {
        cudaEventRecord(e0, st);
        trace_mark(e, "slice start", st);
    }
    if (launch_slice(a, st, e->pdl && e->pre_dirs[0] != nullptr)) return 1;
    // mode 4: the generator is enqueued AFTER the chains (their CTAs are placed first) but only waits for this
    // body's keys, so its small CTAs run next to the chains for the whole slice kernel
    if (e->pre_dirs[0] && e->gen_mode == 4 && enqueue_streams(e, nbuf, st, e->ev_keys)) return 1;
    cudaEventRecord(e1, st);
    trace_mark(e, "slice end", st);
    NSB_LAUNCH_CHECK();
    if (e->pre_dirs[0] && e->gen_mode == 3) {
        // the generator takes the SMs when the chains are done and shares them with the small kernels of step_end
        NSB_CUDA(cudaEventRecord(e->ev_slice, st));
        if (enqueue_streams(e, nbuf, st, e->ev_slice)) return 1;
    }
    e->body += 1;
    e->slice_launches += 1;
    e->all_launches += 3;
    return 0;
}

extern "C" int nsb200_engine_step_end(NsEngine *e, nsb200_stream_t stream) {
    if (!e || !e->initialised) return fail("engine not initialised");
    cudaStream_t st = (cudaStream_t) stream;
    const int D = e->D;
    if (e->p2p) {
        PeerFlags pf;
        for (int r = 0; r < 8; ++r) pf.p[r] = e->peer_flags[r];
        k_peer_barrier<<<1, 32, 0, st>>>(e->ctl, e->p2p_epoch_dev, e->p2p_flags, pf, e->cfg.world_size, e->cfg.rank, e->p2p_err, 0, e->progress_dev);
        e->all_launches += 1;
        trace_mark(e, "peer barrier end", st);
    }
    // the merge overwrites the live buffer whose first m rows the previous body's register update reads
    NSB_CUDA(cudaStreamWaitEvent(st, e->ev_epi[e->slot ^ 1], 0));
    launch_merge_rank(e, e->m, st);
    trace_mark(e, "merge_rank end", st);
    k_merge_scatter<<<592, 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->packed, e->row_doubles, D, e->m, e->N,
                                          (int) e->k, e->rank, e->dead, 1);
    trace_mark(e, "merge_scatter end", st);
    k_iter_advance<<<1, 1, 0, st>>>(e->ctl);
    NSB_CUDA(cudaEventRecord(e->ev_adv, st));
    // register update + loop condition on their own stream: the next body's chains start right away.  Grid form:
    // 8 CTAs on ANY free SMs (a cluster would wait until one GPC has 8 free SMs)
    // sequential runs keep the update on the caller's stream (round 1's schedule: it then runs in the generator's shadow)
    cudaStream_t es = e->no_spec ? st : e->epi_stream;
    if (!e->no_spec) NSB_CUDA(cudaStreamWaitEvent(es, e->ev_adv, 0));
    if (opt(OPT_EPI_CLUSTER) && e->epi_ctas == kEvCluster)
        k_iter_epilogue<<<kEvCluster, kEvThreads, 0, es>>>(e->ctl, e->reg, e->live[0], e->live[1], e->packed,
                                                                      e->row_doubles, D, e->m, e->N, e->tc, 0, e->tabT,
                                                                      e->tabT2, e->tabt, e->N, e->epi, e->progress_dev);
    else
        k_iter_epilogue_grid<<<e->epi_ctas, kEvThreads, 0, es>>>(e->ctl, e->reg, e->live[0], e->live[1], e->packed,
                                                                           e->row_doubles, D, e->m, e->N, e->tc, 0, e->tabT,
                                                                           e->tabT2, e->tabt, e->N, e->epi, e->progress_dev);
    NSB_CUDA(cudaEventRecord(e->ev_epi[e->slot], es));
    NSB_LAUNCH_CHECK();
    trace_mark(e, "  register update end", es);
    if (e->slice_launches == 64) trace_dump();
    e->all_launches += 4;
    return 0;
}

extern "C" int nsb200_engine_step(NsEngine *e, nsb200_stream_t stream) {
    if (!e) return fail("NULL engine");
    if (e->cfg.world_size != 1 && !e->p2p)
        return fail("engine_step requires world_size == 1 or connected peers (nsb200_engine_p2p_connect); otherwise use "
                    "step_begin / all-gather / step_end");
    if (e->external) return fail("family EXTERNAL: drive the body with step_begin / engine_split_* / step_end");
    if (nsb200_engine_step_begin(e, stream)) return 1;
    return nsb200_engine_step_end(e, stream);
}

// ---- engine bodies with a caller-evaluated likelihood ---------------------------------------------
static int engine_split_args(NsEngine *e, SplitArgs &a) {
    if (!e || !e->initialised) return fail("engine not initialised");
    if (!e->external) return fail("engine_split_* requires model.family == NSB200_FAM_EXTERNAL");
    const long long begin = e->rows_per_rank * e->cfg.rank;
    memset(&a, 0, sizeof(a));
    a.model = prior_only(e->cfg.model);
    a.seed_table = e->seed_table;
    a.N = e->N;
    a.chain_begin = begin;
    a.chain_end = begin + e->rows_per_rank;
    a.S = e->cfg.num_slices;
    a.k = e->cfg.num_phantom;
    a.midpoint = e->cfg.midpoint_shrink;
    a.ctl = e->ctl;
    a.live0 = e->live[0];
    a.live1 = e->live[1];
    a.grad_flags = e->grad_flags;
    a.P = e->split_P;
    a.state = split_state_view(e->split_ws, e->D, e->rows_per_rank, (int) e->k);
    return 0;
}

extern "C" int nsb200_engine_set_split_flags(NsEngine *e, int32_t flags) {
    if (!e) return fail("NULL engine");
    if (!e->external) return fail("split flags belong to the split path: create the engine with family EXTERNAL");
    const int P = (flags >> 8) & 0xF;
    if (flags < 0 || (flags & ~0xF03) || P > kSplitMaxP)
        return fail("split flags: bit 0 = gradient_slice, bit 1 = gradient_guided, bits 8-11 = proposals per round (<= %d)", kSplitMaxP);
    e->grad_flags = flags & 3;
    e->split_P = P < 1 ? 1 : P;
    return 0;
}

extern "C" int nsb200_engine_split_grad_points(NsEngine *e, double *out_U, nsb200_stream_t stream) {
    SplitArgs a;
    if (engine_split_args(e, a)) return 1;
    if (!out_U) return fail("out_U is NULL");
    k_split_export_U0<<<296, 256, 0, (cudaStream_t) stream>>>(e->ctl, a.state, e->rows_per_rank, e->D, out_U);
    NSB_LAUNCH_CHECK();
    e->all_launches += 1;
    return 0;
}

extern "C" int nsb200_engine_split_grad_begin(NsEngine *e, const double *grad, double *prop_U, double *prop_X,
                                              uint64_t *n_active, nsb200_stream_t stream) {
    SplitArgs a;
    if (engine_split_args(e, a)) return 1;
    if (!a.grad_flags) return fail("nsb200_engine_set_split_flags with gradient bits first");
    if (!grad || !prop_U) return fail("NULL pointer");
    a.grad = grad;
    a.prop_U = prop_U;
    a.prop_X = prop_X;
    a.active = (unsigned long long *) n_active;
    e->all_launches += 1;
    return launch_split_step(a, 2, (cudaStream_t) stream);
}

extern "C" int nsb200_engine_split_begin(NsEngine *e, double *prop_U, double *prop_X, nsb200_stream_t stream) {
    SplitArgs a;
    if (engine_split_args(e, a)) return 1;
    if (!prop_U) return fail("prop_U is NULL");
    a.prop_U = prop_U;
    a.prop_X = prop_X;
    e->all_launches += 1;
    return launch_split_step(a, 0, (cudaStream_t) stream);
}

extern "C" int nsb200_engine_split_accept(NsEngine *e, const double *prop_logL, double *prop_U, double *prop_X,
                                          uint64_t *n_active, nsb200_stream_t stream) {
    SplitArgs a;
    if (engine_split_args(e, a)) return 1;
    if (!prop_logL || !prop_U) return fail("NULL pointer");
    a.prop_logL = prop_logL;
    a.prop_U = prop_U;
    a.prop_X = prop_X;
    a.active = (unsigned long long *) n_active;
    e->all_launches += 1;
    return launch_split_step(a, 1, (cudaStream_t) stream);
}

extern "C" int nsb200_engine_split_finish(NsEngine *e, nsb200_stream_t stream) {
    SplitArgs a;
    if (engine_split_args(e, a)) return 1;
    k_split_finish<<<296, 256, 0, (cudaStream_t) stream>>>(e->ctl, a.state, e->rows_per_rank, e->D, (int) e->k, nullptr, nullptr,
                                                            nullptr, nullptr, nullptr,
                                                            e->packed + a.chain_begin * e->row_doubles, e->row_doubles);
    NSB_LAUNCH_CHECK();
    e->all_launches += 1;
    return 0;
}

extern "C" int nsb200_engine_contour(NsEngine *e, const double **contour) {
    if (!e || !contour) return fail("NULL pointer");
    *contour = &e->ctl->contour;
    return 0;
}

extern "C" int nsb200_engine_gather_buffer(NsEngine *e, double **buf, int64_t *rows_per_rank, int64_t *row_doubles) {
    if (!e) return fail("NULL engine");
    if (buf) *buf = e->packed;
    if (rows_per_rank) *rows_per_rank = e->rows_per_rank;
    if (row_doubles) *row_doubles = e->row_doubles;
    return 0;
}

extern "C" int nsb200_engine_register(NsEngine *e, NsRegister *out, nsb200_stream_t stream) {
    if (!e || !out) return fail("NULL argument");
    cudaStream_t st = (cudaStream_t) stream;
    NSB_CUDA(cudaStreamSynchronize(st));
    NSB_CUDA(cudaStreamSynchronize(e->epi_stream));  // the register update of the last body runs there
    NSB_CUDA(cudaMemcpyAsync(e->reg_host, e->reg, sizeof(NsRegister), cudaMemcpyDeviceToHost, st));
    NSB_CUDA(cudaStreamSynchronize(st));
    drain_events(e);
    *out = *e->reg_host;
    return 0;
}

extern "C" int nsb200_engine_progress(NsEngine *e, int64_t *completed, int32_t *done) {
    if (!e) return fail("NULL engine");
    const long long c = e->progress[0];
    const long long d = e->progress[1];
    if (completed) *completed = c;
    if (done) *done = (c >= 0 && d) ? 1 : 0;
    return 0;
}

extern "C" int nsb200_engine_finalize(NsEngine *e, nsb200_stream_t stream) {
    if (!e || !e->initialised) return fail("engine not initialised");
    cudaStream_t st = (cudaStream_t) stream;
    NSB_CUDA(cudaStreamWaitEvent(st, e->ev_epi[0], 0));
    NSB_CUDA(cudaStreamWaitEvent(st, e->ev_epi[1], 0));
    // a body that started before the register of its predecessor said "done" is discarded (DevCtl::done_iter)
    k_rollback<<<1, 1, 0, st>>>(e->ctl, e->m, e->k);
    k_blank_rows<<<148, 256, 0, st>>>(e->ctl, e->dead, e->m, e->k, e->D);
    k_append_live<<<296, 256, 0, st>>>(e->ctl, e->live[0], e->live[1], e->dead, e->N, e->D, 1);
    k_finalize_ctl<<<1, 1, 0, st>>>(e->ctl, e->N, e->cap);
    NSB_LAUNCH_CHECK();
    e->all_launches += 4;
    return 0;
}

extern "C" int nsb200_engine_run(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond,
                                 int64_t max_iterations, NsRegister *out_register, nsb200_stream_t stream) {
    if (!e) return fail("NULL engine");
    if (e->cfg.world_size != 1 && !e->p2p) return fail("engine_run requires world_size == 1 or connected peers");
    if (e->external) return fail("family EXTERNAL: the caller drives the loop (engine_init_external / step_begin / engine_split_* / step_end)");
    if (nsb200_engine_init(e, key, term_cond, stream)) return 1;
    // Steps are no-ops on the device once the register says done, so the host runs ahead: it keeps
    // `depth` bodies in flight and polls two host-mapped words the epilogue writes, never the stream.
    const long long depth = opt(OPT_DEPTH) > 0 ? opt(OPT_DEPTH) : 4;
    long long launched = 0, last_completed = -2;
    auto last_progress = std::chrono::steady_clock::now();
    unsigned idle = 0;
    for (;;) {
        const long long completed = e->progress[0];
        if (completed >= 0 && e->progress[1]) break;                                 // loop condition false
        if (max_iterations >= 0 && launched >= max_iterations) {
            if (completed >= launched) break;
        } else if (completed >= 0 ? (launched - completed < depth) : (launched < 1)) {
            if (nsb200_engine_step(e, stream)) return 1;
            ++launched;
            continue;
        }
        // Nothing to enqueue: wait for the device.  Any sticky error (illegal address, ECC, launch timeout, a peer
        // fault) would leave the progress words frozen, so everything but "not ready" is fatal, a rank that missed
        // the all-gather barrier is reported through the third progress word, and a watchdog bounds the wait.
        const cudaError_t q = cudaStreamQuery((cudaStream_t) stream);
        if (q != cudaSuccess && q != cudaErrorNotReady)
            return fail("device error during the run: %s", cudaGetErrorString(q));
        if (e->progress[2]) return fail("a peer GPU did not reach the all-gather barrier (fused NVLink exchange)");
        const auto now = std::chrono::steady_clock::now();
        if (completed != last_completed) {
            last_completed = completed;
            last_progress = now;
        } else if (std::chrono::duration<double>(now - last_progress).count() > 300.0) {
            return fail("no progress on the device for 300 s (body %lld of %lld enqueued)", completed, launched);
        }
        if (++idle > 64) std::this_thread::yield();  // a run spends < 1 ms per body: spin briefly, then be polite
    }
    NsRegister r;
    if (nsb200_engine_finalize(e, stream)) return 1;
    if (nsb200_engine_register(e, &r, stream)) return 1;  // synchronises the stream
    if (out_register) *out_register = r;
    if (r.error_flags & NSB200_ERR_SHRINK_LOOP)
        return fail("a slice chain did not accept within %d proposals: the likelihood is non-deterministic or NaN at "
                    "its seed point (the run was stopped at iteration %lld)", kMaxShrinkProposals, (long long) r.iteration);
    if (r.error_flags & NSB200_ERR_PEER_TIMEOUT) return fail("a peer GPU did not reach the all-gather barrier");
    return 0;
}

extern "C" int nsb200_engine_state(NsEngine *e, NsStateView *out, nsb200_stream_t stream) {
    if (!e || !out) return fail("NULL argument");
    cudaStream_t st = (cudaStream_t) stream;
    NSB_CUDA(cudaStreamSynchronize(e->epi_stream));
    NSB_CUDA(cudaMemcpyAsync(e->ctl_host, e->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost, st));
    NSB_CUDA(cudaStreamSynchronize(st));
    drain_events(e);
    const LiveSet &live = e->live[e->ctl_host->cur];
    out->sender_node_idx = (int64_t *) e->dead.sender;
    out->log_L = e->dead.logL;
    out->U_samples = e->dead.U;
    out->num_likelihood_evaluations = (int64_t *) e->dead.nevals;
    out->phantom = e->dead.phantom;
    out->live_sender_node_idx = (int64_t *) live.sender;
    out->live_U = live.U;
    out->live_log_L_constraint = live.logL_constraint;
    out->live_log_L = live.logL;
    out->live_num_likelihood_evaluations = (int64_t *) live.nevals;
    out->key[0] = e->ctl_host->key.a;
    out->key[1] = e->ctl_host->key.b;
    out->next_sample_idx = e->ctl_host->next_idx;
    out->num_samples = e->ctl_host->num_samples;
    out->capacity = e->cap;
    out->num_live_points = e->N;
    out->D = e->D;
    out->reserved = 0;
    return 0;
}

extern "C" int nsb200_engine_slice_profile(NsEngine *e, double *slice_ms, int64_t *slice_launches, int64_t *all_launches) {
    if (!e) return fail("NULL engine");
    if (slice_ms) *slice_ms = e->slice_ms;
    if (slice_launches) *slice_launches = e->slice_launches;
    if (all_launches) *all_launches = e->all_launches;
    return 0;
}

// -------------------------------------------------------------------------------------------------
// diagnostics: FP64 FMA peak (roofline denominator of the slice kernel)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_fma_peak(double *sink, long long iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0,
           x6 = x0 + 6.0, x7 = x0 + 7.0;
    for (long long i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) sink[0] = s;
}

extern "C" int nsb200_bench_fp64_fma(int64_t iters, double *out_tflops) {
    if (!out_tflops) return fail("out_tflops is NULL");
    if (iters <= 0) iters = 1 << 14;
    int dev = 0, sms = 0;
    NSB_CUDA(cudaGetDevice(&dev));
    NSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    NSB_CUDA(cudaMalloc(&sink, 8));
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, 0);
        k_fp64_fma_peak<<<blocks, threads>>>(sink, iters, 0.999999, 1e-7);
        cudaEventRecord(e1, 0);
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) {
            cudaFree(sink);
            return fail("fp64 peak kernel failed: %s", cudaGetErrorString(err));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double) iters * blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *out_tflops = best;
    return 0;
}

#ifdef NSB_TIMELINE
// reset = 1: arm the stamps for the next generator / slice pair; returns the current stamps first
extern "C" int nsb200_debug_timeline(unsigned long long *out4, int reset) {
    cudaDeviceSynchronize();
    if (out4) cudaMemcpyFromSymbol(out4, nsb::g_tl, 4 * 8);
    if (reset) {
        unsigned long long z[4] = {~0ull, 0ull, ~0ull, 0ull};
        cudaMemcpyToSymbol(nsb::g_tl, z, 4 * 8);
    }
    return 0;
}
#endif

#ifdef NSB_PROFILE
extern "C" int nsb200_debug_profile(unsigned long long *out16, int reset) {
    if (out16) {
        cudaMemcpyFromSymbol(out16, nsb::g_prof, 16 * 8);
        cudaMemcpyFromSymbol(out16 + 9, nsb::g_tail_trips, 8);
    }
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(nsb::g_prof, z, 16 * 8);
        cudaMemcpyToSymbol(nsb::g_tail_trips, z, 8);
    }
    return 0;
}
#endif

/*
 * nsb200.h -- C ABI of the B200-native nested-sampling hot path (drop-in for the static
 * sampler path of Joshuaalbert/jaxns 2.6.9).
 *
 * Plain C: pointers, sizes and POD structs only.  Every device pointer is caller-owned (torch /
 * XLA allocator); kernels are enqueued on the caller's stream and the stateless entry points
 * never allocate or synchronise.  All entry points return 0 on success, non-zero on error with a
 * thread-local message available from nsb200_last_error().  No C++ exception crosses the ABI.
 *
 * The reference has no FFI of its own (it is pure Python on JAX); each entry point cites the
 * reference interface it replaces (paths relative to /root/reference/src/jaxns).  The binding a
 * jaxns maintainer would add (ctypes today, jax.ffi custom calls where JAX is present) is shown in
 * INTEGRATION.md.
 *
 * dtypes follow the reference's policy (internals/mixed_precision.py:88-107):
 *   measure = float64, index = count = int64, num_live_points_per_sample = int32.
 */
#ifndef NSB200_H
#define NSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSB200_ABI_VERSION 2

typedef void *nsb200_stream_t; /* cudaStream_t */

/* ---- registered likelihood families (fused device functions inside the slice kernel) -------- */
enum {
    NSB200_FAM_GAUSS_DENSE = 0,    /* params = [c, mu[D], Linv[D*D] row-major lower]                  */
    NSB200_FAM_GAUSS_MIX_DIAG = 1, /* K x [logc, mean[D], inv_sigma[D]], combined with logaddexp       */
    NSB200_FAM_EGGBOX = 2,         /* (2 + prod_j cos(x_j / 2))^5, no params                           */
    NSB200_FAM_ROSENBROCK = 3,     /* -sum_i 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2, no params          */
    NSB200_FAM_SHELLS = 4,         /* K x [w, r, c[D]] Gaussian shells combined with logaddexp         */
    NSB200_FAM_EXTERNAL = 5        /* likelihood evaluated by the caller on the device between the
                                      propose / accept kernels (nsb200_split_*); no params          */
};

/* ---- prior quantile transforms U in [0,1]^D -> X (framework/wrapped_tfp_distribution.py:77-84) */
enum {
    NSB200_PRIOR_UNIFORM = 0, /* X = U * prior_b + prior_a   (prior_a = low, prior_b = high - low)  */
    NSB200_PRIOR_NORMAL = 1   /* X = ndtri(U) * prior_b + prior_a   (loc, scale)                    */
};

/* Replaces Model.forward / Model.transform (framework/model.py:155-176): the composition
 * U -> quantile -> log_likelihood for a registered family.  All pointers are DEVICE pointers. */
typedef struct NsModelDesc {
    int32_t family;
    int32_t D;
    int32_t prior_kind;
    int32_t K; /* mixture components / shells; 0 otherwise */
    const double *prior_a;
    const double *prior_b;
    const double *params;
    int64_t n_params;
} NsModelDesc;

/* UniDimSliceSampler fields (samplers/uni_slice_sampler.py:296-303) + get_samples arguments
 * (nested_samplers/sharded/sharded_static.py:88). */
typedef struct NsSliceParams {
    int32_t num_slices;      /* S */
    int32_t num_phantom;     /* k (num_phantom_save) */
    int32_t midpoint_shrink; /* bool */
    int32_t split_flags;     /* split path only.  bit 0 = gradient_slice, bit 1 = gradient_guided (:202-214, :255-269);
                              * bits 8-11 = P, proposals per chain and round (0 or 1: one; at most 8): the P proposals are
                              * drawn as if the earlier ones of the round were rejected and evaluated in one batched
                              * likelihood call -- results are identical for every P, the calls per slice drop ~P-fold
                              * until ~1 */
    int64_t num_live;    /* N: rows of live_U / live_logL (sorted ascending) */
    int64_t num_samples; /* m: keys = split(key, m) */
    int64_t chain_begin; /* this GPU evaluates chains [chain_begin, chain_end) -- the */
    int64_t chain_end;   /* PartitionSpec('shard') block of sharded_static.py:104-110 */
} NsSliceParams;

/* TerminationCondition (nested_samplers/common/types.py:26-60).  A field participates iff its
 * bit is set in `mask` (bit i = i-th field in declaration order below). */
typedef struct NsTermCond {
    uint32_t mask;
    uint32_t reserved;
    double ess;                            /* bit 0  */
    double evidence_uncert;                /* bit 1  */
    double live_evidence_frac;             /* bit 2 (deprecated in the reference, ignored) */
    double dlogZ;                          /* bit 3  */
    double max_samples;                    /* bit 4  */
    double max_num_likelihood_evaluations; /* bit 5  */
    double log_L_contour;                  /* bit 6  */
    double efficiency_threshold;           /* bit 7  */
    double rtol;                           /* bit 8  */
    double atol;                           /* bit 9  */
    double peak_XL_frac;                   /* bit 10 */
} NsTermCond;

/* EvidenceCalculation (common/types.py:12-24), field order preserved. */
typedef struct NsEvidenceCalc {
    double log_L, log_X_mean, log_X2_mean, log_Z_mean, log_ZX_mean, log_Z2_mean, log_dZ_mean, log_dZ2_mean;
} NsEvidenceCalc;

/* TerminationRegister (common/types.py:127-138) + the loop's decision. */
typedef struct NsRegister {
    int64_t num_samples_used;
    NsEvidenceCalc evidence_calc;
    NsEvidenceCalc evidence_calc_with_remaining;
    int64_t num_likelihood_evaluations;
    double log_L_contour;
    double efficiency;
    int32_t plateau;
    int32_t no_seed_points;
    double relative_spread;
    double absolute_spread;
    double peak_log_XL;
    int32_t done;               /* determine_termination(...)[0] */
    int32_t error_flags;        /* NSB200_ERR_* bits raised by the device loop (0 = clean run) */
    int64_t termination_reason; /* determine_termination(...)[1], bit map termination.py:17-29 */
    int64_t iteration;
} NsRegister;

/* ShardedStaticNestedSampler fields (sharded_static.py:614-623) after __post_init__ rounding. */
typedef struct NsEngineConfig {
    NsModelDesc model;
    int64_t num_live_points; /* N */
    int64_t max_samples;     /* capacity of the dead-point store (rows) */
    int64_t shell_size;      /* m = int(N * shell_fraction) */
    int32_t num_slices;
    int32_t num_phantom;
    int32_t midpoint_shrink;
    int32_t intended_sender; /* 0 = reference behaviour sender = next_sample_idx - 1 */
    int32_t rank;            /* chains [rank*m/world, (rank+1)*m/world) are evaluated locally */
    int32_t world_size;
} NsEngineConfig;

/* Device-pointer view of NestedSamplerState + LivePointCollection (common/types.py:108-148). */
typedef struct NsStateView {
    /* SampleCollection, capacity max_samples */
    int64_t *sender_node_idx;
    double *log_L;
    double *U_samples; /* [max_samples, D] */
    int64_t *num_likelihood_evaluations;
    uint8_t *phantom;
    /* LivePointCollection, N rows, sorted ascending by log_L */
    int64_t *live_sender_node_idx;
    double *live_U; /* [N, D] */
    double *live_log_L_constraint;
    double *live_log_L;
    int64_t *live_num_likelihood_evaluations;
    /* scalars (host copies, valid after nsb200_engine_sync) */
    uint32_t key[2];
    int64_t next_sample_idx;
    int64_t num_samples;
    int64_t capacity;
    int64_t num_live_points;
    int32_t D;
    int32_t reserved;
} NsStateView;

typedef struct NsEngine NsEngine;

/* Device-side error bits (NsRegister.error_flags; the whole-run entry points turn them into an error return). */
enum {
    NSB200_ERR_SHRINK_LOOP = 1, /* a slice did not accept within 65536 proposals: the likelihood is non-deterministic
                                   or NaN at its own seed point (the reference's lax.while_loop would spin forever,
                                   samplers/uni_slice_sampler.py:160-196) */
    NSB200_ERR_PEER_TIMEOUT = 2, /* a peer GPU did not reach the fused all-gather's arrival barrier */
    NSB200_ERR_CONTOUR_MISMATCH = 4 /* the ranks' likelihood contours (L_min), exchanged with the arrival flags of the
                                       fused all-gather, disagree: their replicated live sets have diverged */
};

/* ---- misc ----------------------------------------------------------------------------------- */
int nsb200_abi_version(void);
const char *nsb200_last_error(void);
/* Tuning / A-B knobs (launch geometry, kernel selection: results never depend on them).  Every knob is also an
 * environment variable of the same name, read once per process; value < 0 returns to that default.  Names:
 * NSB200_SPEC, NSB200_TPB, NSB200_SLICE_MMA, NSB200_MMA_P, NSB200_MMA_WPB, NSB200_MERGE_BRUTE, NSB200_GEN_MODE,
 * NSB200_GEN_SMS, NSB200_GEN_TPB, NSB200_GEN_FENCE, NSB200_SPECULATE, NSB200_EPI_CLUSTER, NSB200_EPI_CTAS,
 * NSB200_EPI_PRIO, NSB200_DEPTH, NSB200_TRACE. */
int nsb200_set_option(const char *name, int32_t value);

/* Copies the two words of a PRNG key from DEVICE memory (where XLA keeps keys) to the host; synchronises `stream`.
 * For bindings that receive keys as device operands (csrc/xla_ffi_shim.cc): the entry points take keys by value. */
int nsb200_read_key(const uint32_t *device_key, uint32_t out[2], nsb200_stream_t stream);

/* ---- jax.random under jax_threefry_partitionable=True (internals/mixed_precision.py:11-15) --- */
/* threefry2x32 primitive over n counter pairs (device arrays). */
int nsb200_threefry2x32(const uint32_t key[2], const uint32_t *x0, const uint32_t *x1, int64_t n,
                        uint32_t *out0, uint32_t *out1, nsb200_stream_t stream);
/* jax.random.split(key, n) -> out[n,2] (device).  Call sites: sharded_static.py:128,248,491,510. */
int nsb200_random_split(const uint32_t key[2], int64_t n, uint32_t *out, nsb200_stream_t stream);
/* jax.random.bits(key, (n,), uint64) */
int nsb200_random_bits64(const uint32_t key[2], int64_t n, uint64_t *out, nsb200_stream_t stream);
/* jax.random.uniform(key, (n,), float64, lo, hi)  (model.py:136, uni_slice_sampler.py:83) */
int nsb200_random_uniform(const uint32_t key[2], int64_t n, double lo, double hi, double *out,
                          nsb200_stream_t stream);
/* jax.random.normal(key, (n,), float64)  (uni_slice_sampler.py:36) */
int nsb200_random_normal(const uint32_t key[2], int64_t n, double *out, nsb200_stream_t stream);

/* ---- model ---------------------------------------------------------------------------------- */
/* vmap(Model.forward)(U) and optionally vmap(Model.transform)(U)  (framework/model.py:155-176).
 * U [n,D]; out_logL [n]; out_X [n,D] or NULL. */
int nsb200_forward_batch(const NsModelDesc *model, const double *U, int64_t n, double *out_logL,
                         double *out_X, nsb200_stream_t stream);

/* ---- B1: batched constrained samplers (get_samples, sharded_static.py:88-129) ---------------- */
/* Seed-choice table c[q] = logaddexp-cumsum of q+1 zeros: what cumulative_logsumexp
 * (internals/log_semiring.py:51-92) yields on the suffix mask of random.py:78-84.  out [N] device. */
int nsb200_seed_table(int64_t N, double *out, nsb200_stream_t stream);

/* draw_uniform_samples over keys split(sample_key, N)[begin:end]
 * (common/initialisation.py:47-60, common/uniform_sample.py:12-60). */
int nsb200_init_batch(const NsModelDesc *model, const uint32_t sample_key[2], int64_t N, int64_t begin,
                      int64_t end, double *out_U, double *out_logL, int64_t *out_nevals,
                      nsb200_stream_t stream);

/* get_samples with UniDimSliceSampler(perfect=True): chains [chain_begin, chain_end) of
 * split(key, m); each chain = BaseAbstractMarkovSampler._get_sample (samplers/bases.py:63-75).
 * `contour` is a DEVICE pointer to log_L_contour (so the loop never round-trips to the host).
 * live_U [N,D], live_logL [N] ascending, seed_table [N].
 * out_U [n,D], out_logL [n], out_nevals [n], ph_U [n*k,D], ph_logL [n*k] with n = chain_end-chain_begin
 * (ph_* may be NULL when k == 0). */
int nsb200_slice_batch(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                       const double *contour, const double *live_U, const double *live_logL,
                       const double *seed_table, double *out_U, double *out_logL, int64_t *out_nevals,
                       double *ph_U, double *ph_logL, nsb200_stream_t stream);

/* Same call with caller-provided scratch for the chains' data-independent streams (directions, proposal
 * uniforms, continuation keys: nothing in a chain's key tree depends on a likelihood value, so they are produced
 * by a throughput kernel first).  With the streams available the dense-Gaussian family with D <= 32 runs on the
 * FP64 tensor-core path (8 proposals per warp through mma.sync.m8n8k4.f64); results are identical to
 * nsb200_slice_batch up to the summation order of the quadratic form.  workspace_bytes >=
 * nsb200_slice_streams_bytes(D, num_slices, chain_end - chain_begin). */
int64_t nsb200_slice_streams_bytes(int32_t D, int32_t num_slices, int64_t n_chains);
int nsb200_slice_batch_ws(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                          const double *contour, const double *live_U, const double *live_logL,
                          const double *seed_table, double *out_U, double *out_logL, int64_t *out_nevals,
                          double *ph_U, double *ph_logL, void *workspace, int64_t workspace_bytes,
                          int32_t *error_flags, nsb200_stream_t stream);

/* ---- B1 for likelihoods the library cannot fuse: the slice step split around a caller-evaluated
 * batched likelihood (BASELINE north_star).  Same chains as nsb200_slice_batch
 * (samplers/bases.py:63-75, samplers/uni_slice_sampler.py:114-273,343-441) when the likelihood
 * values agree.  Protocol, all on one stream, n = chain_end - chain_begin:
 *     nsb200_split_begin(...)            -> prop_U / prop_X [P,n,D]: first proposals of every chain (P = proposals per
 *                                           round from NsSliceParams.split_flags, default 1; row p*n + i = proposal p
 *                                           of chain i)
 *     repeat:  prop_logL[P*n] = vmap(log_likelihood)(prop_X)      (caller; XLA / torch on the device)
 *              nsb200_split_accept(...)  -> accept or shrink (NaN -> -inf, framework/ops.py:323-325),
 *                                           next proposal, *n_active += chains still running
 *     until n_active == 0, then nsb200_split_finish(...) -> the Sample / phantom arrays.
 * `model` supplies the prior transform only (family is ignored, NSB200_FAM_EXTERNAL allowed).
 * Finished chains keep their final point in prop_U (their prop_logL is ignored), so shapes are
 * static.  `workspace` holds the chain state between calls. */
int64_t nsb200_split_workspace_bytes(int32_t D, int64_t n_chains, int32_t num_phantom);
int nsb200_split_begin(const NsModelDesc *model, const NsSliceParams *p, const uint32_t key[2],
                       const double *contour, const double *live_U, const double *live_logL,
                       const double *seed_table, void *workspace, int64_t workspace_bytes, double *prop_U,
                       double *prop_X, nsb200_stream_t stream);
/* n_active: DEVICE counter (uint64) incremented by every chain that still needs evaluations, or NULL.  Bit 62 is set
 * when a chain was stopped by the shrink-loop watchdog (NSB200_ERR_SHRINK_LOOP: 65536 proposals of one slice
 * rejected -- a non-deterministic likelihood, or NaN at the chain's own seed point). */
int nsb200_split_accept(const NsModelDesc *model, const NsSliceParams *p, const double *contour,
                        const double *prop_logL, void *workspace, int64_t workspace_bytes, double *prop_U,
                        double *prop_X, uint64_t *n_active, nsb200_stream_t stream);
int nsb200_split_finish(const NsModelDesc *model, const NsSliceParams *p, void *workspace,
                        int64_t workspace_bytes, double *out_U, double *out_logL, int64_t *out_nevals,
                        double *ph_U, double *ph_logL, nsb200_stream_t stream);
/* Gradient variants of UniDimSliceSampler (samplers/uni_slice_sampler.py:202-214 gradient_slice: the slice direction is
 * the normalised gradient of log L w.r.t. U at the chain's point and only the uphill half of the bracket is searched;
 * :255-269 gradient_guided: the next direction is the Householder reflection of the last one about the gradient at the
 * accepted point).  With gradient bits in NsSliceParams.split_flags a chain that is about to start a slice waits for the
 * caller's gradient (jax.grad in the reference; any autodiff of the batched likelihood here):
 *     nsb200_split_begin, then per round:
 *         nsb200_split_grad_points -> U0 [n,D]; caller: grad [n,D] = d log L / dU at U0
 *         nsb200_split_grad_begin  -> waiting chains start their slice (first proposals into prop_U / prop_X)
 *         caller: prop_logL;  nsb200_split_accept  (accepting chains wait again)
 * Every gradient counts as one likelihood evaluation, as in the reference. */
int nsb200_split_grad_points(const NsModelDesc *model, const NsSliceParams *p, void *workspace, int64_t workspace_bytes,
                             double *out_U, nsb200_stream_t stream);
int nsb200_split_grad_begin(const NsModelDesc *model, const NsSliceParams *p, const double *contour, const double *grad,
                            void *workspace, int64_t workspace_bytes, double *prop_U, double *prop_X, uint64_t *n_active,
                            nsb200_stream_t stream);
/* Redraw round `round` (0 = first draw) of _single_uniform_sample (common/uniform_sample.py:12-60) for
 * prior draws [begin, end) of split(sample_key, N); rows with need[i - begin] == 0 are left untouched
 * (need == NULL: all rows).  out_U / out_X [end - begin, D] (out_X optional). */
int nsb200_init_propose(const NsModelDesc *model, const uint32_t sample_key[2], int64_t N, int64_t begin,
                        int64_t end, int32_t round, const uint8_t *need, double *out_U, double *out_X,
                        nsb200_stream_t stream);
/* vmap(Model.transform)(U) alone (framework/model.py:155-159): U [n,D] -> X [n,D]. */
int nsb200_transform_batch(const NsModelDesc *model, const double *U, int64_t n, double *out_X,
                           nsb200_stream_t stream);

/* get_samples with UniformSampler (samplers/uniform_samplers.py:42-85, max_likelihood_evals=100). */
int nsb200_uniform_batch(const NsModelDesc *model, const uint32_t key[2], const double *contour,
                         int64_t num_samples, int64_t chain_begin, int64_t chain_end, double *out_U,
                         double *out_logL, int64_t *out_nevals, nsb200_stream_t stream);

/* ---- statistics over the dead-point set ----------------------------------------------------- */
enum { NSB200_WS_ARGSORT = 0, NSB200_WS_COUNT_CROSSED_EDGES = 1, NSB200_WS_EVIDENCE_STATS = 2,
       NSB200_WS_LOGSUMEXP = 3 };
/* Scratch bytes the caller must provide for an op over n elements. */
int64_t nsb200_workspace_bytes(int32_t op, int64_t n);

/* jnp.argsort(keys) (stable; -0 == +0; NaN last) -> out_idx int64 [n]
 * (sharded_static.py:174,274; initialisation.py:75). */
int nsb200_argsort_f64(const double *keys, int64_t n, int64_t *out_idx, void *workspace,
                       int64_t workspace_bytes, nsb200_stream_t stream);

/* count_crossed_edges (internals/tree_structure.py:33-108).  sender int64 [M], log_L [M];
 * num_samples < 0 selects the static variant (trim=True), otherwise the dynamic variant.
 * out_samples_indices int64 [M], out_num_live_points int32 [M]. */
int nsb200_count_crossed_edges(const int64_t *sender_node_idx, const double *log_L, int64_t M,
                               int64_t num_samples, int64_t *out_samples_indices,
                               int32_t *out_num_live_points, void *workspace, int64_t workspace_bytes,
                               nsb200_stream_t stream);

/* compute_evidence_stats / cumulative_op_static(_update_evidence_calc_op)
 * (internals/shrinkage_statistics.py:43-157).  init: HOST pointer to the starting
 * EvidenceCalculation (NULL = create_init_evidence_calc).  log_L [M] sorted, num_live_points
 * float64 [M].  out_final: DEVICE NsEvidenceCalc.  out_per_sample: DEVICE [8, M] field-major in
 * NsEvidenceCalc order, or NULL. */
int nsb200_evidence_stats(const NsEvidenceCalc *init, const double *log_L, const double *num_live_points,
                          int64_t M, NsEvidenceCalc *out_final, double *out_per_sample, void *workspace,
                          int64_t workspace_bytes, nsb200_stream_t stream);

/* LogSpace.sum (internals/log_semiring.py:187-190): logsumexp of x[n] -> out[0] (device). */
int nsb200_logsumexp(const double *x, int64_t n, double *out, void *workspace, int64_t workspace_bytes,
                     nsb200_stream_t stream);

/* ---- multi-GPU: fused all-gather over NVLink peer memory --------------------------------------
 * Replaces the all-gather of the new samples after the sharded get_samples
 * (nested_samplers/sharded/sharded_static.py:104-129, out_specs=PartitionSpec()): once every rank's engine is
 * connected, the slice kernel stores each packed row straight into the gather buffer of every rank and
 * engine_step_end starts with an arrival barrier on device flags -- the host issues no collective.
 * export: the CUDA IPC handle of this engine's arena + byte offsets of {gather buffer 0, gather buffer 1, flags};
 * connect: handles [world][64] and offsets [world][3] of all ranks in rank order (exchange them with any host
 * transport).  enabled(e, -1) queries nothing and is a no-op, 0/1 switch the mode (the host all-gather of
 * nsb200_engine_gather_buffer stays available).  error: 1 if a peer did not arrive within ~2 s. */
int nsb200_engine_p2p_export(NsEngine *e, uint8_t handle[64], int64_t offsets[3]);
int nsb200_engine_p2p_connect(NsEngine *e, const uint8_t *handles, const int64_t *offsets);
int nsb200_engine_p2p_enabled(NsEngine *e, int32_t enable);
int nsb200_engine_p2p_error(NsEngine *e, int32_t *out);

/* sample_evidence (utils.py:433-476): S simulations of the shrinkage over the M dead points, simulation s with
 * split(key, S)[s] and per-point keys split(key_s, M)[i]; log T_i = log(uniform(key_i, ())) / n_i.
 * num_live_points float64 [M], log_L [M] (sorted), out [S] = samples of log Z.  All DEVICE pointers. */
int nsb200_sample_evidence(const uint32_t key[2], const double *num_live_points, const double *log_L, int64_t M,
                           int64_t S, double *out, nsb200_stream_t stream);

/* ---- B2/B3: engine-owned state (ShardedStaticNestedSampler._run, sharded_static.py:775-851) -- */
int nsb200_engine_create(const NsEngineConfig *cfg, NsEngine **out);
void nsb200_engine_destroy(NsEngine *e);
/* create_init_state + create_init_termination_register (common/initialisation.py:20-108). */
int nsb200_engine_init(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond,
                       nsb200_stream_t stream);
/* One loop body of _main_ns_thread (sharded_static.py:487-555) = key splits + _collect_shell
 * (:210-324) + termination decision.  With world_size > 1 the step is split so the host can
 * all-gather the packed new rows between the halves:
 *   step_begin : discard shell, append to the dead store, run this rank's chains -> local rows
 *   step_end   : merge + stable sort, phantom append, register update, termination decision. */
int nsb200_engine_step(NsEngine *e, nsb200_stream_t stream);
int nsb200_engine_step_begin(NsEngine *e, nsb200_stream_t stream);
int nsb200_engine_step_end(NsEngine *e, nsb200_stream_t stream);
/* Packed rows produced by step_begin, laid out for an all-gather:
 * gather buffer = world_size blocks of [rows_per_rank, row_doubles] float64; this rank writes block
 * `rank`.  row = [U[D], log_L, n_evals (int64 bits), (U[D], log_L) x k phantom]. */
int nsb200_engine_gather_buffer(NsEngine *e, double **buf, int64_t *rows_per_rank, int64_t *row_doubles);
/* Device address of the likelihood contour L_min of the body in flight (written by step_begin): with the host-issued
 * collective the caller agrees it across ranks by ncclAllReduce(min) / (max) next to the all-gather
 * (the fused path compares the contours inside its arrival barrier). */
int nsb200_engine_contour(NsEngine *e, const double **contour);
/* Whole run: loops step until determine_termination (common/termination.py:13-147) says done,
 * then appends the final live set (sharded_static.py:834-838).  Synchronises `stream`. */
int nsb200_engine_run(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond,
                      int64_t max_iterations, NsRegister *out_register, nsb200_stream_t stream);
/* Engine with model.family == NSB200_FAM_EXTERNAL: the caller evaluates every likelihood.
 *   init:  U/logL/nevals [N] = the initial live points (nsb200_init_propose rounds + caller's likelihood),
 *          every rank passes the same N rows (create_init_state, common/initialisation.py:20-84);
 *   body:  nsb200_engine_step_begin (discard + append only), nsb200_engine_split_begin, the
 *          likelihood / nsb200_engine_split_accept rounds until *n_active == 0, nsb200_engine_split_finish
 *          (packs this rank's rows into the gather buffer), [all-gather], nsb200_engine_step_end. */
int nsb200_engine_init_external(NsEngine *e, const uint32_t key[2], const NsTermCond *term_cond,
                                const double *U, const double *log_L, const int64_t *num_likelihood_evaluations,
                                nsb200_stream_t stream);
int nsb200_engine_split_begin(NsEngine *e, double *prop_U, double *prop_X, nsb200_stream_t stream);
int nsb200_engine_split_accept(NsEngine *e, const double *prop_logL, double *prop_U, double *prop_X,
                               uint64_t *n_active, nsb200_stream_t stream);
int nsb200_engine_split_finish(NsEngine *e, nsb200_stream_t stream);
/* The same gradient protocol for an engine with model.family == NSB200_FAM_EXTERNAL (flags as NsSliceParams.split_flags). */
int nsb200_engine_set_split_flags(NsEngine *e, int32_t flags);
int nsb200_engine_split_grad_points(NsEngine *e, double *out_U, nsb200_stream_t stream);
int nsb200_engine_split_grad_begin(NsEngine *e, const double *grad, double *prop_U, double *prop_X, uint64_t *n_active,
                                   nsb200_stream_t stream);
/* Final live-set append (sharded_static.py:834-838). */
int nsb200_engine_finalize(NsEngine *e, nsb200_stream_t stream);
/* Non-blocking progress of the enqueued bodies: *completed = bodies whose register update has run on the
 * device (-1 before the init pass finished), *done = loop condition false.  Lets a host-driven loop keep a
 * few bodies in flight without synchronising the stream. */
int nsb200_engine_progress(NsEngine *e, int64_t *completed, int32_t *done);
/* Blocks on `stream`, then copies the register / fills the view. */
int nsb200_engine_register(NsEngine *e, NsRegister *out, nsb200_stream_t stream);
int nsb200_engine_state(NsEngine *e, NsStateView *out, nsb200_stream_t stream);
/* Device time (ms, CUDA events on the launching stream) and launch count of the slice kernel
 * accumulated since engine_init; used by bench.py for the roofline numerator/denominator. */
int nsb200_engine_slice_profile(NsEngine *e, double *slice_ms, int64_t *slice_launches, int64_t *all_launches);

/* ---- diagnostics ----------------------------------------------------------------------------- */
/* Measures the FP64 FMA peak of the current device (independent DFMA chains, CUDA-event timed):
 * the roofline denominator of the fused slice kernel (MEASURED_PEAKS.json has no FP64 figure).
 * Synchronous; returns TFLOP/s in *out_tflops. */
int nsb200_bench_fp64_fma(int64_t iters, double *out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* NSB200_H */

// XLA FFI custom-call handlers over the C ABI of include/nsb200.h: the binding BASELINE.json's north star names
// ("Python host code calls the kernels through thin jax.ffi (XLA FFI C-ABI) custom calls").
//
// Each handler unpacks XLA buffers / attributes and forwards to ONE nsb200_* entry point on the stream XLA hands
// over; buffer order = the C function's argument order, so the reference-side Python is a one-line
// jax.ffi.ffi_call per function.  Registration on the jaxns side:
//     jax.ffi.register_ffi_target("nsb200_slice_batch", jax.ffi.pycapsule(lib.nsb200_ffi_slice_batch), platform="CUDA")
//
// Built only where XLA's header-only FFI API exists (jaxlib ships it: jax.ffi.include_dir()):
//     g++ -std=c++17 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude
//         jaxns_b200/csrc/xla_ffi_shim.cc -Ljaxns_b200 -lnsb200 -o jaxns_b200/libnsb200_ffi.so
// This image has no jax; tests/test_host_cpu.py::test_xla_ffi_shim_compiles compiles the file against the stand-in
// header csrc/ffi_stub/xla/ffi/api/ffi.h, which type-checks every handler against its binding.
//
// Replaces (reference call sites): get_samples -> sampler.get_sample (sharded_static.py:88-129), draw_uniform_samples
// (common/uniform_sample.py:63), vmap(Model.forward) (framework/model.py:167-176), count_crossed_edges
// (internals/tree_structure.py:33-108), compute_evidence_stats (internals/shrinkage_statistics.py:131-157),
// sample_evidence (utils.py:433-476).
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>
#include <string>

#include "xla/ffi/api/ffi.h"

#include "../../include/nsb200.h"

namespace ffi = xla::ffi;
using Stream = void *;  // cudaStream_t; the shim itself needs no CUDA header

static ffi::Error Status(int rc) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error::Internal(std::string("nsb200: ") + nsb200_last_error());
}

// XLA hands PRNG keys over as device buffers; the C ABI takes the two key words by value (they seed a launch, they
// are not data).  nsb200_read_key copies them off the device on the call's stream (one 8-byte D2H + stream sync).
static ffi::Error ReadKey(const ffi::Buffer<ffi::U32> &key, uint32_t out[2], Stream stream) {
    if (key.element_count() != 2) return ffi::Error::InvalidArgument("key must be uint32[2]");
    return Status(nsb200_read_key(key.typed_data(), out, stream));
}

static NsModelDesc ModelDesc(int32_t family, int32_t prior_kind, int32_t K, const ffi::Buffer<ffi::F64> &prior_a,
                             const ffi::Buffer<ffi::F64> &prior_b, const ffi::Buffer<ffi::F64> &params) {
    NsModelDesc m;
    m.family = family;
    m.D = (int32_t) prior_a.element_count();
    m.prior_kind = prior_kind;
    m.K = K;
    m.prior_a = prior_a.typed_data();
    m.prior_b = prior_b.typed_data();
    m.params = params.typed_data();
    m.n_params = (int64_t) params.element_count();
    return m;
}

// ---- get_samples with UniDimSliceSampler: chains [chain_begin, chain_end) of split(key, num_samples) ------------
static ffi::Error SliceBatchImpl(Stream stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F64> contour,
                                 ffi::Buffer<ffi::F64> live_U, ffi::Buffer<ffi::F64> live_logL,
                                 ffi::Buffer<ffi::F64> seed_table, ffi::Buffer<ffi::F64> prior_a,
                                 ffi::Buffer<ffi::F64> prior_b, ffi::Buffer<ffi::F64> params,
                                 ffi::ResultBuffer<ffi::F64> out_U, ffi::ResultBuffer<ffi::F64> out_logL,
                                 ffi::ResultBuffer<ffi::S64> out_nevals, ffi::ResultBuffer<ffi::F64> ph_U,
                                 ffi::ResultBuffer<ffi::F64> ph_logL, ffi::ResultBuffer<ffi::U8> workspace,
                                 ffi::ResultBuffer<ffi::S32> error_flags, int32_t family, int32_t prior_kind, int32_t K,
                                 int32_t num_slices, int32_t num_phantom, int32_t midpoint_shrink, int64_t num_samples,
                                 int64_t chain_begin, int64_t chain_end) {
    if (live_U.dimensions().size() != 2) return ffi::Error::InvalidArgument("live_U must be [N, D]");
    uint32_t k2[2];
    if (ffi::Error err = ReadKey(key, k2, stream); !err.success()) return err;
    const NsModelDesc m = ModelDesc(family, prior_kind, K, prior_a, prior_b, params);
    NsSliceParams p;
    p.num_slices = num_slices;
    p.num_phantom = num_phantom;
    p.midpoint_shrink = midpoint_shrink;
    p.split_flags = 0;
    p.num_live = live_U.dimensions()[0];
    p.num_samples = num_samples;
    p.chain_begin = chain_begin;
    p.chain_end = chain_end;
    return Status(nsb200_slice_batch_ws(&m, &p, k2, contour.typed_data(), live_U.typed_data(), live_logL.typed_data(),
                                        seed_table.typed_data(), out_U->typed_data(), out_logL->typed_data(),
                                        out_nevals->typed_data(), num_phantom ? ph_U->typed_data() : nullptr,
                                        num_phantom ? ph_logL->typed_data() : nullptr, workspace->typed_data(),
                                        (int64_t) workspace->element_count(), error_flags->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_slice_batch, SliceBatchImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<int32_t>("family")
                                  .Attr<int32_t>("prior_kind")
                                  .Attr<int32_t>("K")
                                  .Attr<int32_t>("num_slices")
                                  .Attr<int32_t>("num_phantom")
                                  .Attr<int32_t>("midpoint_shrink")
                                  .Attr<int64_t>("num_samples")
                                  .Attr<int64_t>("chain_begin")
                                  .Attr<int64_t>("chain_end"));

// ---- draw_uniform_samples over split(sample_key, N)[begin:end] --------------------------------------------------
static ffi::Error InitBatchImpl(Stream stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F64> prior_a,
                                ffi::Buffer<ffi::F64> prior_b, ffi::Buffer<ffi::F64> params,
                                ffi::ResultBuffer<ffi::F64> out_U, ffi::ResultBuffer<ffi::F64> out_logL,
                                ffi::ResultBuffer<ffi::S64> out_nevals, int32_t family, int32_t prior_kind, int32_t K,
                                int64_t N, int64_t begin, int64_t end) {
    uint32_t k2[2];
    if (ffi::Error err = ReadKey(key, k2, stream); !err.success()) return err;
    const NsModelDesc m = ModelDesc(family, prior_kind, K, prior_a, prior_b, params);
    return Status(nsb200_init_batch(&m, k2, N, begin, end, out_U->typed_data(), out_logL->typed_data(),
                                    out_nevals->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_init_batch, InitBatchImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Attr<int32_t>("family")
                                  .Attr<int32_t>("prior_kind")
                                  .Attr<int32_t>("K")
                                  .Attr<int64_t>("N")
                                  .Attr<int64_t>("begin")
                                  .Attr<int64_t>("end"));

// ---- vmap(Model.forward) (+ vmap(Model.transform)) --------------------------------------------------------------
static ffi::Error ForwardBatchImpl(Stream stream, ffi::Buffer<ffi::F64> U, ffi::Buffer<ffi::F64> prior_a,
                                   ffi::Buffer<ffi::F64> prior_b, ffi::Buffer<ffi::F64> params,
                                   ffi::ResultBuffer<ffi::F64> out_logL, ffi::ResultBuffer<ffi::F64> out_X,
                                   int32_t family, int32_t prior_kind, int32_t K) {
    if (U.dimensions().size() != 2) return ffi::Error::InvalidArgument("U must be [n, D]");
    const NsModelDesc m = ModelDesc(family, prior_kind, K, prior_a, prior_b, params);
    return Status(nsb200_forward_batch(&m, U.typed_data(), U.dimensions()[0], out_logL->typed_data(),
                                       out_X->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_forward_batch, ForwardBatchImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("family")
                                  .Attr<int32_t>("prior_kind")
                                  .Attr<int32_t>("K"));

// ---- count_crossed_edges ----------------------------------------------------------------------------------------
static ffi::Error CountCrossedEdgesImpl(Stream stream, ffi::Buffer<ffi::S64> sender, ffi::Buffer<ffi::F64> log_L,
                                        ffi::ResultBuffer<ffi::S64> samples_indices,
                                        ffi::ResultBuffer<ffi::S32> num_live_points, ffi::ResultBuffer<ffi::U8> workspace,
                                        int64_t num_samples) {
    const int64_t M = (int64_t) log_L.element_count();
    if ((int64_t) sender.element_count() != M) return ffi::Error::InvalidArgument("sender and log_L differ in length");
    return Status(nsb200_count_crossed_edges(sender.typed_data(), log_L.typed_data(), M, num_samples,
                                             samples_indices->typed_data(), num_live_points->typed_data(),
                                             workspace->typed_data(), (int64_t) workspace->element_count(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_count_crossed_edges, CountCrossedEdgesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("num_samples"));

// ---- compute_evidence_stats: out_final = 8 doubles in EvidenceCalculation order, per_sample [8, M] ------------------
static ffi::Error EvidenceStatsImpl(Stream stream, ffi::Buffer<ffi::F64> log_L, ffi::Buffer<ffi::F64> num_live_points,
                                    ffi::ResultBuffer<ffi::F64> out_final, ffi::ResultBuffer<ffi::F64> per_sample,
                                    ffi::ResultBuffer<ffi::U8> workspace) {
    const int64_t M = (int64_t) log_L.element_count();
    if (out_final->element_count() != 8) return ffi::Error::InvalidArgument("out_final must be float64[8]");
    return Status(nsb200_evidence_stats(nullptr, log_L.typed_data(), num_live_points.typed_data(), M,
                                        reinterpret_cast<NsEvidenceCalc *>(out_final->typed_data()),
                                        per_sample->typed_data(), workspace->typed_data(),
                                        (int64_t) workspace->element_count(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_evidence_stats, EvidenceStatsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// ---- sample_evidence: S simulations of the shrinkage -------------------------------------------------------------
static ffi::Error SampleEvidenceImpl(Stream stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F64> num_live_points,
                                     ffi::Buffer<ffi::F64> log_L, ffi::ResultBuffer<ffi::F64> out) {
    uint32_t k2[2];
    if (ffi::Error err = ReadKey(key, k2, stream); !err.success()) return err;
    return Status(nsb200_sample_evidence(k2, num_live_points.typed_data(), log_L.typed_data(),
                                         (int64_t) log_L.element_count(), (int64_t) out->element_count(),
                                         out->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_sample_evidence, SampleEvidenceImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// ---- arbitrary JAX likelihoods: the slice step split around vmap(log_likelihood) (include/nsb200.h nsb200_split_*) -----
// XLA buffers are values, the chain state is not: `workspace`, `prop_U` (the accepted point of a round is read back from
// it) and the `n_active` counter are operands AND results of the same call and must be bound with
// jax.ffi.ffi_call(..., input_output_aliases={...}) so that both name one allocation (the handlers refuse otherwise;
// `n_active` comes in as zeros).  One jitted round on the jaxns side is then
//     logL = jax.vmap(model.log_prob_likelihood)(prop_U)                       # the user's JAX likelihood, on the device
//     ws, prop_U, prop_X, n_active = ffi_call("nsb200_split_accept", ...)(contour, logL, a, b, ws, prop_U, zeros)
// inside a lax.while_loop on n_active > 0 (uni_slice_sampler.py:114-273 is the loop this replaces).
static NsModelDesc ExternalDesc(int32_t prior_kind, const ffi::Buffer<ffi::F64> &prior_a, const ffi::Buffer<ffi::F64> &prior_b) {
    NsModelDesc m;
    m.family = NSB200_FAM_EXTERNAL;
    m.D = (int32_t) prior_a.element_count();
    m.prior_kind = prior_kind;
    m.K = 0;
    m.prior_a = prior_a.typed_data();
    m.prior_b = prior_b.typed_data();
    m.params = nullptr;
    m.n_params = 0;
    return m;
}

static NsSliceParams SplitParams(int32_t num_slices, int32_t num_phantom, int32_t midpoint_shrink, int32_t split_flags,
                                 int64_t num_live, int64_t num_samples, int64_t chain_begin, int64_t chain_end) {
    NsSliceParams p;
    p.num_slices = num_slices;
    p.num_phantom = num_phantom;
    p.midpoint_shrink = midpoint_shrink;
    p.split_flags = split_flags;
    p.num_live = num_live;
    p.num_samples = num_samples;
    p.chain_begin = chain_begin;
    p.chain_end = chain_end;
    return p;
}

static ffi::Error SplitBeginImpl(Stream stream, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F64> contour,
                                 ffi::Buffer<ffi::F64> live_U, ffi::Buffer<ffi::F64> live_logL,
                                 ffi::Buffer<ffi::F64> seed_table, ffi::Buffer<ffi::F64> prior_a,
                                 ffi::Buffer<ffi::F64> prior_b, ffi::ResultBuffer<ffi::U8> workspace,
                                 ffi::ResultBuffer<ffi::F64> prop_U, ffi::ResultBuffer<ffi::F64> prop_X,
                                 int32_t prior_kind, int32_t num_slices, int32_t num_phantom, int32_t midpoint_shrink,
                                 int32_t split_flags, int64_t num_samples, int64_t chain_begin, int64_t chain_end) {
    if (live_U.dimensions().size() != 2) return ffi::Error::InvalidArgument("live_U must be [N, D]");
    uint32_t k2[2];
    if (ffi::Error err = ReadKey(key, k2, stream); !err.success()) return err;
    const NsModelDesc m = ExternalDesc(prior_kind, prior_a, prior_b);
    const NsSliceParams p = SplitParams(num_slices, num_phantom, midpoint_shrink, split_flags, live_U.dimensions()[0],
                                        num_samples, chain_begin, chain_end);
    return Status(nsb200_split_begin(&m, &p, k2, contour.typed_data(), live_U.typed_data(), live_logL.typed_data(),
                                     seed_table.typed_data(), workspace->typed_data(),
                                     (int64_t) workspace->element_count(), prop_U->typed_data(), prop_X->typed_data(),
                                     stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_split_begin, SplitBeginImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("prior_kind")
                                  .Attr<int32_t>("num_slices")
                                  .Attr<int32_t>("num_phantom")
                                  .Attr<int32_t>("midpoint_shrink")
                                  .Attr<int32_t>("split_flags")
                                  .Attr<int64_t>("num_samples")
                                  .Attr<int64_t>("chain_begin")
                                  .Attr<int64_t>("chain_end"));

static ffi::Error SplitAcceptImpl(Stream stream, ffi::Buffer<ffi::F64> contour, ffi::Buffer<ffi::F64> prop_logL,
                                  ffi::Buffer<ffi::F64> prior_a, ffi::Buffer<ffi::F64> prior_b,
                                  ffi::Buffer<ffi::U8> workspace_in, ffi::Buffer<ffi::F64> prop_U_in,
                                  ffi::Buffer<ffi::U64> n_active_in, ffi::ResultBuffer<ffi::U8> workspace,
                                  ffi::ResultBuffer<ffi::F64> prop_U, ffi::ResultBuffer<ffi::F64> prop_X,
                                  ffi::ResultBuffer<ffi::U64> n_active, int32_t prior_kind, int32_t num_slices,
                                  int32_t num_phantom, int32_t midpoint_shrink, int32_t split_flags, int64_t num_live,
                                  int64_t num_samples, int64_t chain_begin, int64_t chain_end) {
    if (workspace_in.typed_data() != workspace->typed_data() || prop_U_in.typed_data() != prop_U->typed_data() ||
        n_active_in.typed_data() != n_active->typed_data())
        return ffi::Error::InvalidArgument("workspace, prop_U and n_active must be bound with input_output_aliases");
    const NsModelDesc m = ExternalDesc(prior_kind, prior_a, prior_b);
    const NsSliceParams p = SplitParams(num_slices, num_phantom, midpoint_shrink, split_flags, num_live, num_samples,
                                        chain_begin, chain_end);
    return Status(nsb200_split_accept(&m, &p, contour.typed_data(), prop_logL.typed_data(), workspace->typed_data(),
                                      (int64_t) workspace->element_count(), prop_U->typed_data(), prop_X->typed_data(),
                                      n_active->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_split_accept, SplitAcceptImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U64>>()
                                  .Attr<int32_t>("prior_kind")
                                  .Attr<int32_t>("num_slices")
                                  .Attr<int32_t>("num_phantom")
                                  .Attr<int32_t>("midpoint_shrink")
                                  .Attr<int32_t>("split_flags")
                                  .Attr<int64_t>("num_live")
                                  .Attr<int64_t>("num_samples")
                                  .Attr<int64_t>("chain_begin")
                                  .Attr<int64_t>("chain_end"));

static ffi::Error SplitFinishImpl(Stream stream, ffi::Buffer<ffi::F64> prior_a, ffi::Buffer<ffi::F64> prior_b,
                                  ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F64> out_U,
                                  ffi::ResultBuffer<ffi::F64> out_logL, ffi::ResultBuffer<ffi::S64> out_nevals,
                                  ffi::ResultBuffer<ffi::F64> ph_U, ffi::ResultBuffer<ffi::F64> ph_logL,
                                  int32_t prior_kind, int32_t num_slices, int32_t num_phantom, int32_t midpoint_shrink,
                                  int32_t split_flags, int64_t num_live, int64_t num_samples, int64_t chain_begin,
                                  int64_t chain_end) {
    const NsModelDesc m = ExternalDesc(prior_kind, prior_a, prior_b);
    const NsSliceParams p = SplitParams(num_slices, num_phantom, midpoint_shrink, split_flags, num_live, num_samples,
                                        chain_begin, chain_end);
    // the C entry point only reads the workspace; XLA hands operands over as const
    return Status(nsb200_split_finish(&m, &p, const_cast<uint8_t *>(workspace.typed_data()),
                                      (int64_t) workspace.element_count(), out_U->typed_data(), out_logL->typed_data(),
                                      out_nevals->typed_data(), num_phantom ? ph_U->typed_data() : nullptr,
                                      num_phantom ? ph_logL->typed_data() : nullptr, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(nsb200_ffi_split_finish, SplitFinishImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<Stream>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("prior_kind")
                                  .Attr<int32_t>("num_slices")
                                  .Attr<int32_t>("num_phantom")
                                  .Attr<int32_t>("midpoint_shrink")
                                  .Attr<int32_t>("split_flags")
                                  .Attr<int64_t>("num_live")
                                  .Attr<int64_t>("num_samples")
                                  .Attr<int64_t>("chain_begin")
                                  .Attr<int64_t>("chain_end"));

#endif  // __has_include("xla/ffi/api/ffi.h")

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

#endif  // __has_include("xla/ffi/api/ffi.h")

"""
ctypes binding of the C ABI declared in include/nsb200.h (libnsb200.so, built by csrc/build.sh or
__graft_entry__.build()).  The product path has no CPU fallback: if the shared library is missing or
CUDA is unavailable, every entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NSB200_LIB: explicit path of the shared library (kernel experiments: A/B builds side by side)
_SO = os.environ.get("NSB200_LIB") or os.path.join(_HERE, "libnsb200.so")
_lib = None

c_f64p = ctypes.c_void_p
c_stream = ctypes.c_void_p


class NsModelDesc(ctypes.Structure):
    _fields_ = [("family", ctypes.c_int32), ("D", ctypes.c_int32), ("prior_kind", ctypes.c_int32),
                ("K", ctypes.c_int32), ("prior_a", ctypes.c_void_p), ("prior_b", ctypes.c_void_p),
                ("params", ctypes.c_void_p), ("n_params", ctypes.c_int64)]


class NsSliceParams(ctypes.Structure):
    _fields_ = [("num_slices", ctypes.c_int32), ("num_phantom", ctypes.c_int32),
                ("midpoint_shrink", ctypes.c_int32), ("split_flags", ctypes.c_int32),
                ("num_live", ctypes.c_int64), ("num_samples", ctypes.c_int64),
                ("chain_begin", ctypes.c_int64), ("chain_end", ctypes.c_int64)]


class NsTermCond(ctypes.Structure):
    _fields_ = [("mask", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("ess", ctypes.c_double),
                ("evidence_uncert", ctypes.c_double), ("live_evidence_frac", ctypes.c_double),
                ("dlogZ", ctypes.c_double), ("max_samples", ctypes.c_double),
                ("max_num_likelihood_evaluations", ctypes.c_double), ("log_L_contour", ctypes.c_double),
                ("efficiency_threshold", ctypes.c_double), ("rtol", ctypes.c_double), ("atol", ctypes.c_double),
                ("peak_XL_frac", ctypes.c_double)]


TERM_FIELDS = ("ess", "evidence_uncert", "live_evidence_frac", "dlogZ", "max_samples",
               "max_num_likelihood_evaluations", "log_L_contour", "efficiency_threshold", "rtol", "atol",
               "peak_XL_frac")


class NsEvidenceCalc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("log_L", "log_X_mean", "log_X2_mean", "log_Z_mean", "log_ZX_mean", "log_Z2_mean", "log_dZ_mean",
                 "log_dZ2_mean")]


class NsRegister(ctypes.Structure):
    _fields_ = [("num_samples_used", ctypes.c_int64), ("evidence_calc", NsEvidenceCalc),
                ("evidence_calc_with_remaining", NsEvidenceCalc), ("num_likelihood_evaluations", ctypes.c_int64),
                ("log_L_contour", ctypes.c_double), ("efficiency", ctypes.c_double), ("plateau", ctypes.c_int32),
                ("no_seed_points", ctypes.c_int32), ("relative_spread", ctypes.c_double),
                ("absolute_spread", ctypes.c_double), ("peak_log_XL", ctypes.c_double), ("done", ctypes.c_int32),
                ("error_flags", ctypes.c_int32), ("termination_reason", ctypes.c_int64), ("iteration", ctypes.c_int64)]


class NsEngineConfig(ctypes.Structure):
    _fields_ = [("model", NsModelDesc), ("num_live_points", ctypes.c_int64), ("max_samples", ctypes.c_int64),
                ("shell_size", ctypes.c_int64), ("num_slices", ctypes.c_int32), ("num_phantom", ctypes.c_int32),
                ("midpoint_shrink", ctypes.c_int32), ("intended_sender", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("world_size", ctypes.c_int32)]


class NsStateView(ctypes.Structure):
    _fields_ = [("sender_node_idx", ctypes.c_void_p), ("log_L", ctypes.c_void_p), ("U_samples", ctypes.c_void_p),
                ("num_likelihood_evaluations", ctypes.c_void_p), ("phantom", ctypes.c_void_p),
                ("live_sender_node_idx", ctypes.c_void_p), ("live_U", ctypes.c_void_p),
                ("live_log_L_constraint", ctypes.c_void_p), ("live_log_L", ctypes.c_void_p),
                ("live_num_likelihood_evaluations", ctypes.c_void_p), ("key", ctypes.c_uint32 * 2),
                ("next_sample_idx", ctypes.c_int64), ("num_samples", ctypes.c_int64), ("capacity", ctypes.c_int64),
                ("num_live_points", ctypes.c_int64), ("D", ctypes.c_int32), ("reserved", ctypes.c_int32)]


WS_ARGSORT, WS_COUNT_CROSSED_EDGES, WS_EVIDENCE_STATS, WS_LOGSUMEXP = 0, 1, 2, 3

# every symbol include/nsb200.h declares
EXPORTS = (
    "nsb200_abi_version", "nsb200_last_error", "nsb200_set_option", "nsb200_read_key", "nsb200_threefry2x32", "nsb200_random_split",
    "nsb200_random_bits64", "nsb200_random_uniform", "nsb200_random_normal", "nsb200_forward_batch",
    "nsb200_seed_table", "nsb200_init_batch", "nsb200_slice_batch", "nsb200_uniform_batch",
    "nsb200_workspace_bytes", "nsb200_argsort_f64", "nsb200_count_crossed_edges", "nsb200_evidence_stats",
    "nsb200_logsumexp", "nsb200_engine_create", "nsb200_engine_destroy", "nsb200_engine_init",
    "nsb200_engine_step", "nsb200_engine_step_begin", "nsb200_engine_step_end", "nsb200_engine_gather_buffer",
    "nsb200_engine_run", "nsb200_engine_finalize", "nsb200_engine_register", "nsb200_engine_state",
    "nsb200_engine_slice_profile", "nsb200_bench_fp64_fma", "nsb200_engine_progress",
    "nsb200_split_workspace_bytes", "nsb200_split_begin", "nsb200_split_accept", "nsb200_split_finish",
    "nsb200_init_propose", "nsb200_transform_batch", "nsb200_engine_init_external", "nsb200_engine_split_begin",
    "nsb200_engine_split_accept", "nsb200_engine_split_finish", "nsb200_sample_evidence",
    "nsb200_slice_streams_bytes", "nsb200_slice_batch_ws", "nsb200_split_grad_points", "nsb200_split_grad_begin",
    "nsb200_engine_set_split_flags", "nsb200_engine_contour", "nsb200_engine_split_grad_points", "nsb200_engine_split_grad_begin",
    "nsb200_engine_p2p_export", "nsb200_engine_p2p_connect", "nsb200_engine_p2p_enabled", "nsb200_engine_p2p_error",
)


def so_path() -> str:
    return _SO


def lib():
    """Load libnsb200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                f"g.build()' or jaxns_b200/csrc/build.sh). jaxns_b200 has no CPU fallback.")
        L = ctypes.CDLL(_SO)
        L.nsb200_last_error.restype = ctypes.c_char_p
        L.nsb200_workspace_bytes.restype = ctypes.c_int64
        L.nsb200_workspace_bytes.argtypes = [ctypes.c_int32, ctypes.c_int64]
        L.nsb200_engine_destroy.restype = None
        L.nsb200_split_workspace_bytes.restype = ctypes.c_int64
        L.nsb200_split_workspace_bytes.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32]
        L.nsb200_slice_streams_bytes.restype = ctypes.c_int64
        L.nsb200_slice_streams_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64]
        _lib = L
    return _lib


def set_option(name: str, value: int):
    """Tuning / A-B knob of the library (include/nsb200.h nsb200_set_option); value < 0 restores the default."""
    check(lib().nsb200_set_option(name.encode(), ctypes.c_int32(int(value))))


def check(rc: int):
    if rc != 0:
        raise RuntimeError("nsb200: " + lib().nsb200_last_error().decode())


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("jaxns_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")


def key_arg(key):
    """uint32[2] host array from a PRNGKey-like (sequence / numpy / torch)."""
    import numpy as np
    try:
        import torch
        if isinstance(key, torch.Tensor):
            key = key.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    k = np.asarray(key).astype(np.uint32).reshape(2)
    return (ctypes.c_uint32 * 2)(int(k[0]), int(k[1]))


def stream_arg(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)

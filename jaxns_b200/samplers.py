"""
Sampler plugin interface and the two samplers of the static loop, with the reference's fields and
validation (/root/reference/src/jaxns/samplers/abc.py:10-115, samplers/bases.py:12-75,
samplers/uni_slice_sampler.py:277-341, samplers/uniform_samplers.py:21-41).  The batched entry point
`get_samples_batch` is what get_samples (nested_samplers/sharded/sharded_static.py:88-129) calls.
"""
import ctypes
import dataclasses
import warnings
from typing import Any, NamedTuple, Tuple

import torch

from jaxns_b200 import _lib
from jaxns_b200.types import LivePointCollection, Sample, TerminationRegister

__all__ = ["UniDimSliceSampler", "UniformSampler", "AbstractSampler", "EphemeralState", "SeedPoint"]

# Proposals per chain and round of the split (caller-evaluated likelihood) path: each batched likelihood call evaluates
# this many speculative proposals per chain (include/nsb200.h NsSliceParams.split_flags).  Results do not depend on it;
# the likelihood calls per slice drop from ~3.8 (1) to ~1.3 (4) and ~1.1 (8).  1..8; 4 measured best on the config-2
# problem with a torch likelihood (profiles/r2/external_path_r2.txt).
SPLIT_PROPOSALS = 4


class EphemeralState(NamedTuple):
    key: Any
    live_points_collection: LivePointCollection
    termination_register: TerminationRegister


class SeedPoint(NamedTuple):
    U0: Any
    log_L0: Any


class AbstractSampler:
    def num_phantom(self) -> int:
        return 0

    def pre_process(self, ephemeral_state: EphemeralState) -> Any:
        return self._pre_process(ephemeral_state)

    def post_process(self, ephemeral_state: EphemeralState, sampler_state: Any) -> Any:
        return self._post_process(ephemeral_state, sampler_state)

    def _pre_process(self, ephemeral_state):
        return ephemeral_state.live_points_collection

    def _post_process(self, ephemeral_state, sampler_state):
        return ephemeral_state.live_points_collection

    def get_samples_batch(self, key, log_L_constraint, sampler_state, num_samples: int, chain_begin: int = 0,
                          chain_end: int = None) -> Tuple[Sample, Sample]:
        raise NotImplementedError


def _contour_tensor(log_L_constraint) -> torch.Tensor:
    if isinstance(log_L_constraint, torch.Tensor):
        return log_L_constraint.to(device="cuda", dtype=torch.float64).reshape(1)
    return torch.tensor([float(log_L_constraint)], dtype=torch.float64, device="cuda")


@dataclasses.dataclass(eq=False)
class UniDimSliceSampler(AbstractSampler):
    model: Any
    num_slices: int
    num_phantom_save: int
    midpoint_shrink: bool
    perfect: bool
    gradient_slice: bool = False
    adaptive_shrink: bool = False
    gradient_guided: bool = False

    def __post_init__(self):
        if self.num_slices < 1:
            raise ValueError(f"num_slices should be >= 1, got {self.num_slices}.")
        if self.num_phantom_save < 0:
            raise ValueError(f"num_phantom_save should be >= 0, got {self.num_phantom_save}.")
        if self.num_phantom_save >= self.num_slices:
            raise ValueError(
                f"num_phantom_save should be < num_slices, got {self.num_phantom_save} >= {self.num_slices}.")
        self.num_slices = int(self.num_slices)
        self.num_phantom_save = int(self.num_phantom_save)
        self.midpoint_shrink = bool(self.midpoint_shrink)
        self.perfect = bool(self.perfect)
        self.gradient_slice = bool(self.gradient_slice)
        self.adaptive_shrink = bool(self.adaptive_shrink)
        self.gradient_guided = bool(self.gradient_guided)
        if self.adaptive_shrink:
            raise NotImplementedError("Adaptive shrinkage not implemented.")
        if not self.perfect:
            raise ValueError("Only perfect slice sampler is implemented.")
        if self.gradient_guided:
            warnings.warn("Gradient guided slice sampler is experimental and will likely change.")
        self._seed_tables = {}

    def num_phantom(self) -> int:
        return self.num_phantom_save

    @property
    def gradient_flags(self) -> int:
        """NsSliceParams.gradient_flags: bit 0 gradient_slice, bit 1 gradient_guided.  Non-zero: the chains run through
        the split propose / accept step with the model's gradient supplied between the kernels (Model.grad_U)."""
        return int(self.gradient_slice) | (int(self.gradient_guided) << 1)

    @property
    def split_proposals(self) -> int:
        return int(getattr(self, "_split_proposals", None) or SPLIT_PROPOSALS)

    @split_proposals.setter
    def split_proposals(self, P: int):
        if not 1 <= int(P) <= 8:
            raise ValueError("split_proposals must be in 1..8")
        self._split_proposals = int(P)

    @property
    def split_flags(self) -> int:
        return self.gradient_flags | (self.split_proposals << 8)

    def _seed_table(self, N: int) -> torch.Tensor:
        t = self._seed_tables.get(N)
        if t is None:
            t = torch.empty(N, dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib().nsb200_seed_table(ctypes.c_int64(N), _lib.ptr(t), _lib.stream_arg()))
            self._seed_tables[N] = t
        return t

    def get_samples_batch(self, key, log_L_constraint, sampler_state: LivePointCollection, num_samples: int,
                          chain_begin: int = 0, chain_end: int = None) -> Tuple[Sample, Sample]:
        """Chains [chain_begin, chain_end) of split(key, num_samples), each one
        BaseAbstractMarkovSampler._get_sample (samplers/bases.py:63-75)."""
        _lib.require_cuda()
        if chain_end is None:
            chain_end = num_samples
        live_U = sampler_state.U_sample.contiguous()
        live_logL = sampler_state.log_L.contiguous()
        N, D = live_U.shape
        n = chain_end - chain_begin
        k = self.num_phantom_save
        contour = _contour_tensor(log_L_constraint)
        out_U = torch.empty((n, D), dtype=torch.float64, device="cuda")
        out_logL = torch.empty(n, dtype=torch.float64, device="cuda")
        out_nev = torch.empty(n, dtype=torch.int64, device="cuda")
        ph_U = torch.empty((n * k, D), dtype=torch.float64, device="cuda")
        ph_logL = torch.empty(n * k, dtype=torch.float64, device="cuda")
        p = _lib.NsSliceParams(self.num_slices, k, int(self.midpoint_shrink), self.split_flags, N, int(num_samples),
                               int(chain_begin), int(chain_end))
        d = self.model.desc(external=bool(self.gradient_flags))
        if getattr(self.model, "is_external", False) or self.gradient_flags:
            self._split_batch(d, p, key, contour, live_U, live_logL, out_U, out_logL, out_nev, ph_U, ph_logL)
        else:
            self._fused_batch(d, p, key, contour, live_U, live_logL, out_U, out_logL, out_nev, ph_U, ph_logL)
        cons = contour.expand(n)
        sample = Sample(U_sample=out_U, log_L_constraint=cons, log_L=out_logL, num_likelihood_evaluations=out_nev)
        phantom = Sample(U_sample=ph_U, log_L_constraint=contour.expand(n * k), log_L=ph_logL,
                         num_likelihood_evaluations=torch.zeros(n * k, dtype=torch.int64, device="cuda"))
        return sample, phantom

    def _split_batch(self, d, p, key, contour, live_U, live_logL, out_U, out_logL, out_nev, ph_U, ph_logL):
        """The slice step split around the model's batched device likelihood (include/nsb200.h nsb200_split_*)."""
        L = _lib.lib()
        n, D = out_U.shape
        k = self.num_phantom_save
        N = live_U.shape[0]
        if n == 0:
            return
        nbytes = L.nsb200_split_workspace_bytes(D, n, k)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        P = self.split_proposals
        prop_U = torch.full((P * n, D), 0.5, dtype=torch.float64, device="cuda")  # [P, n, D]
        prop_X = torch.zeros((P * n, D), dtype=torch.float64, device="cuda")
        active = torch.zeros(1, dtype=torch.int64, device="cuda")
        st = _lib.stream_arg()
        _lib.check(L.nsb200_split_begin(ctypes.byref(d), ctypes.byref(p), _lib.key_arg(key), _lib.ptr(contour),
                                        _lib.ptr(live_U), _lib.ptr(live_logL), _lib.ptr(self._seed_table(N)),
                                        _lib.ptr(ws), ctypes.c_int64(nbytes), _lib.ptr(prop_U), _lib.ptr(prop_X), st))
        burst = max(4, self.num_slices // (4 * min(P, 4)))  # likelihood rounds between reads of the active-chain counter
        grad_pts = torch.empty((n, D), dtype=torch.float64, device="cuda") if self.gradient_flags else None
        while True:
            for r in range(burst):
                if grad_pts is not None:
                    # chains between slices wait for d log L / dU at their point (include/nsb200.h, gradient protocol)
                    _lib.check(L.nsb200_split_grad_points(ctypes.byref(d), ctypes.byref(p), _lib.ptr(ws),
                                                          ctypes.c_int64(nbytes), _lib.ptr(grad_pts), st))
                    grad = self.model.grad_U(grad_pts)
                    _lib.check(L.nsb200_split_grad_begin(ctypes.byref(d), ctypes.byref(p), _lib.ptr(contour),
                                                         _lib.ptr(grad), _lib.ptr(ws), ctypes.c_int64(nbytes),
                                                         _lib.ptr(prop_U), _lib.ptr(prop_X), ctypes.c_void_p(0), st))
                logL = self.model.external_log_likelihood(prop_U, prop_X)
                last = r == burst - 1
                if last:
                    active.zero_()
                _lib.check(L.nsb200_split_accept(ctypes.byref(d), ctypes.byref(p), _lib.ptr(contour), _lib.ptr(logL),
                                                 _lib.ptr(ws), ctypes.c_int64(nbytes), _lib.ptr(prop_U),
                                                 _lib.ptr(prop_X), _lib.ptr(active) if last else ctypes.c_void_p(0),
                                                 st))
            n_active = int(active.item())
            if n_active & (1 << 62):
                raise RuntimeError("nsb200: a slice chain did not accept within 65536 proposals: the likelihood is "
                                   "non-deterministic or NaN at its seed point")
            if n_active == 0:
                break
        _lib.check(L.nsb200_split_finish(ctypes.byref(d), ctypes.byref(p), _lib.ptr(ws), ctypes.c_int64(nbytes),
                                         _lib.ptr(out_U), _lib.ptr(out_logL), _lib.ptr(out_nev),
                                         _lib.ptr(ph_U) if k else ctypes.c_void_p(0),
                                         _lib.ptr(ph_logL) if k else ctypes.c_void_p(0), st))

    def _fused_batch(self, d, p, key, contour, live_U, live_logL, out_U, out_logL, out_nev, ph_U, ph_logL):
        """nsb200_slice_batch_ws: the chains' data-independent streams are generated into a scratch buffer first, so
        the dense-Gaussian family with D <= 32 runs on the FP64 tensor-core kernel, exactly as inside the engine."""
        L = _lib.lib()
        k = self.num_phantom_save
        n, D = out_U.shape
        N = live_U.shape[0]
        if n == 0:
            return
        nbytes = int(L.nsb200_slice_streams_bytes(D, self.num_slices, n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(L.nsb200_slice_batch_ws(
            ctypes.byref(d), ctypes.byref(p), _lib.key_arg(key), _lib.ptr(contour), _lib.ptr(live_U),
            _lib.ptr(live_logL), _lib.ptr(self._seed_table(N)), _lib.ptr(out_U), _lib.ptr(out_logL),
            _lib.ptr(out_nev), _lib.ptr(ph_U) if k else ctypes.c_void_p(0),
            _lib.ptr(ph_logL) if k else ctypes.c_void_p(0), _lib.ptr(ws), ctypes.c_int64(nbytes), _lib.ptr(err),
            _lib.stream_arg()))
        flags = int(err.item())
        if flags:
            raise RuntimeError(f"nsb200: slice chains raised error flags {flags} (1 = a slice did not accept within "
                               "65536 proposals: non-deterministic or NaN likelihood)")


@dataclasses.dataclass(eq=False)
class UniformSampler(AbstractSampler):
    model: Any
    max_likelihood_evals: int = 100

    def __post_init__(self):
        if self.max_likelihood_evals != 100:
            raise NotImplementedError("The fused uniform sampler is built for max_likelihood_evals=100.")

    def get_samples_batch(self, key, log_L_constraint, sampler_state, num_samples: int, chain_begin: int = 0,
                          chain_end: int = None) -> Tuple[Sample, Sample]:
        _lib.require_cuda()
        if chain_end is None:
            chain_end = num_samples
        n = chain_end - chain_begin
        D = self.model.U_ndims
        contour = _contour_tensor(log_L_constraint)
        out_U = torch.empty((n, D), dtype=torch.float64, device="cuda")
        out_logL = torch.empty(n, dtype=torch.float64, device="cuda")
        out_nev = torch.empty(n, dtype=torch.int64, device="cuda")
        d = self.model.desc()
        _lib.check(_lib.lib().nsb200_uniform_batch(ctypes.byref(d), _lib.key_arg(key), _lib.ptr(contour),
                                                    ctypes.c_int64(num_samples), ctypes.c_int64(chain_begin),
                                                    ctypes.c_int64(chain_end), _lib.ptr(out_U), _lib.ptr(out_logL),
                                                    _lib.ptr(out_nev), _lib.stream_arg()))
        sample = Sample(U_sample=out_U, log_L_constraint=contour.expand(n), log_L=out_logL,
                        num_likelihood_evaluations=out_nev)
        empty = Sample(U_sample=out_U[:0], log_L_constraint=out_logL[:0], log_L=out_logL[:0],
                       num_likelihood_evaluations=out_nev[:0])
        return sample, empty

"""
ShardedStaticNestedSampler: the static nested-sampling engine behind the reference's interface
(/root/reference/src/jaxns/nested_samplers/sharded/sharded_static.py:577-851), driving the
C-ABI engine (include/nsb200.h, nsb200_engine_*).  One process per GPU; with torch.distributed
initialised (world_size > 1) the m chains of an iteration are split into contiguous blocks per rank
exactly like PartitionSpec('shard') on split(sample_key, m) (:104-110,128) and the packed new rows
are all-gathered over NCCL between the two halves of a step.
"""
import ctypes
import dataclasses
import math
import os
import sys
import time
import warnings
from typing import Any, List, Optional, Tuple

import numpy as np
import torch

from jaxns_b200 import _lib, termination
from jaxns_b200.internals.shrinkage_statistics import compute_evidence_stats, logsumexp
from jaxns_b200.internals.stats import effective_sample_size_kish, linear_to_log_stats
from jaxns_b200.internals.tree_structure import SampleTreeGraph, count_crossed_edges
from jaxns_b200.samplers import AbstractSampler, UniDimSliceSampler
from jaxns_b200.types import (NestedSamplerResults, NestedSamplerState, Sample, SampleCollection,
                              TerminationCondition, TerminationRegister)

__all__ = ["ShardedStaticNestedSampler"]


def round_up_num_live_points(init_num_live_points, shell_frac, num_devices):
    """sharded_static.py:577-584"""
    num_live_points = int(init_num_live_points)
    while True:
        shell_size = int(num_live_points * shell_frac)
        if shell_size % num_devices == 0:
            break
        num_live_points += 1
    return num_live_points


def round_up_max_samples(init_max_samples, num_discard, num_phantom_points):
    """sharded_static.py:587-594"""
    max_samples = int(init_max_samples)
    block_size = num_discard * (1 + num_phantom_points)
    while True:
        if max_samples % block_size == 0:
            break
        max_samples += 1
    return max_samples


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def get_samples(key, mesh, sampler: AbstractSampler, sampler_state: Any, log_L_contour, num_samples: int
                ) -> Tuple[Sample, Sample]:
    """get_samples (sharded_static.py:88-129).  `mesh` = (rank, world_size) or None; each rank
    evaluates its contiguous block of chains and the blocks are all-gathered."""
    rank, world = mesh if mesh is not None else _dist_info()
    per = num_samples // world
    sample, phantom = sampler.get_samples_batch(key, log_L_contour, sampler_state, num_samples,
                                                chain_begin=rank * per, chain_end=(rank + 1) * per)
    if world == 1:
        return sample, phantom
    import torch.distributed as dist

    def gather(x):
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous())
        return out

    return Sample(*[gather(x) for x in sample]), Sample(*[gather(x) for x in phantom])


class ContourAgreement:
    """(min, max) all-reduce of a replicated scalar next to the per-body all-gather; `check` raises on the ranks that
    differ from a peer.  `contour` is a 1-element float64 tensor that the engine rewrites in place every body."""

    def __init__(self, contour: torch.Tensor, rank: int = 0):
        self.contour = contour
        self.rank = rank
        self.bad = torch.zeros(1, dtype=torch.bool, device=contour.device)

    def all_reduce(self):
        import torch.distributed as dist
        pair = torch.cat([self.contour, -self.contour])
        dist.all_reduce(pair, op=dist.ReduceOp.MIN)
        self.bad |= (pair[0] != self.contour[0]) | (pair[1] != -self.contour[0])

    def check(self):
        if bool(self.bad.item()):
            raise RuntimeError(f"nsb200: rank {self.rank} disagrees with its peers on the likelihood contour L_min "
                               "(the replicated live sets have diverged)")


class _DevView:
    """Zero-copy torch view of engine-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr,
                                             data=(int(ptr), False), version=2)
        self._owner = owner


def _view(ptr, shape, typestr, owner) -> torch.Tensor:
    if int(np.prod(shape)) == 0:
        dt = {"<f8": torch.float64, "<i8": torch.int64, "|u1": torch.uint8}[typestr]
        return torch.empty(tuple(shape), dtype=dt, device="cuda")
    return torch.as_tensor(_DevView(ptr, shape, typestr, owner), device="cuda")


class _Engine:
    """RAII wrapper of NsEngine*."""

    def __init__(self, cfg: _lib.NsEngineConfig, keepalive):
        self._h = ctypes.c_void_p()
        self._keepalive = keepalive
        _lib.check(_lib.lib().nsb200_engine_create(ctypes.byref(cfg), ctypes.byref(self._h)))

    @property
    def h(self):
        return self._h

    def __del__(self):
        try:
            if self._h:
                _lib.lib().nsb200_engine_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:  # interpreter shutdown
            pass


@dataclasses.dataclass(eq=False)
class ShardedStaticNestedSampler:
    model: Any
    max_samples: int
    init_efficiency_threshold: float
    sampler: AbstractSampler
    num_live_points: int
    shell_fraction: Optional[float] = None
    num_dynamic_refinement_iterations: int = 0
    refine_threshold: float = 0.01
    devices: Optional[List[Any]] = None
    verbose: bool = False

    def __post_init__(self):
        if self.shell_fraction is None:
            self.shell_fraction = 0.5
        self.shell_fraction = max(self.shell_fraction, 1. / self.num_live_points)
        if (self.shell_fraction <= 0.) or (self.shell_fraction > 1.):
            raise ValueError(
                f"Expected 0 < shell_fraction <= 1, got {self.shell_fraction}. Best to keep it around 0.5.")
        self._rank, self._world = _dist_info()
        if self.devices is None:
            self.devices = list(range(self._world))
        if len(self.devices) > 1 and self._rank == 0:
            import sys
            print(f"Running over {len(self.devices)} devices.", file=sys.stderr)
        self.num_live_points = round_up_num_live_points(
            init_num_live_points=self.num_live_points,
            shell_frac=self.shell_fraction,
            num_devices=len(self.devices)
        )
        self.max_samples = round_up_max_samples(
            init_max_samples=self.max_samples,
            num_discard=int(self.shell_fraction * self.num_live_points),
            num_phantom_points=self.sampler.num_phantom()
        )
        if self.num_dynamic_refinement_iterations > 0:
            raise NotImplementedError("Dynamic refinement is experimental in the reference and out of scope here.")
        if not isinstance(self.sampler, UniDimSliceSampler):
            raise NotImplementedError("The device loop runs UniDimSliceSampler chains.")
        self._engine = None
        self.last_register = None
        self.last_profile = None

    # ------------------------------------------------------------------------------------------
    def _make_engine(self) -> _Engine:
        if self._engine is None:
            s = self.sampler
            desc = self.model.desc(external=bool(s.gradient_flags))
            eng_world = len(self.devices) if self._world > 1 else 1
            cfg = _lib.NsEngineConfig(desc, int(self.num_live_points), int(self.max_samples),
                                      int(self.num_live_points * self.shell_fraction), s.num_slices,
                                      s.num_phantom_save, int(s.midpoint_shrink), 0,
                                      self._rank if eng_world > 1 else 0, eng_world)
            self._engine = _Engine(cfg, keepalive=(self.model, desc))
            if getattr(self.model, "is_external", False) or s.gradient_flags:
                _lib.check(_lib.lib().nsb200_engine_set_split_flags(self._engine.h, ctypes.c_int32(s.split_flags)))
        return self._engine

    def _run(self, key, term_cond) -> Tuple[int, TerminationRegister, NestedSamplerState]:
        """_run (sharded_static.py:775-851): init, (no-op) uniform phase, slice phase, final append."""
        _lib.require_cuda()
        L = _lib.lib()
        t_dbg = time.perf_counter()
        eng = self._make_engine()
        if os.environ.get("NSB200_LOOP_DEBUG"):
            print(f"[loop rank {self._rank}] engine create {1e3 * (time.perf_counter() - t_dbg):.1f} ms", file=sys.stderr)
        stream = _lib.stream_arg()
        plain = isinstance(term_cond, TerminationCondition)
        if plain:
            if term_cond.live_evidence_frac is not None:
                warnings.warn("live_evidence_frac is deprecated, use dlogZ instead.")
            tc = termination.to_c(term_cond)
        else:
            tc = _lib.NsTermCond()  # device stops only on plateau / no seed points; host decides the rest
        reg = _lib.NsRegister()
        world = len(self.devices) if self._world > 1 else 1
        # gradient_slice / gradient_guided chains take the caller-evaluated path whatever the likelihood is
        external = getattr(self.model, "is_external", False) or bool(self.sampler.gradient_flags)
        # fused NVLink all-gather: with connected peers a body needs no host-issued collective, so
        # the whole loop runs inside the library exactly as on one GPU
        p2p = world > 1 and not external and self._connect_peers(eng, world)
        if external:
            host_tc = self._effective_host_cond(term_cond) if not plain else None
            self._run_external(eng, key, tc, reg, stream, world, host_tc)
        elif plain and (world == 1 or p2p) and not self.verbose:
            _lib.check(L.nsb200_engine_run(eng.h, _lib.key_arg(key), ctypes.byref(tc), ctypes.c_int64(-1),
                                           ctypes.byref(reg), stream))
        elif plain and (world == 1 or p2p):
            # verbose: one body at a time, the register read back and printed after each (sharded_static.py:519-549)
            _lib.check(L.nsb200_engine_init(eng.h, _lib.key_arg(key), ctypes.byref(tc), stream))
            _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
            while not reg.done:
                _lib.check(L.nsb200_engine_step(eng.h, stream))
                _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
                if self._rank == 0:
                    self._print_register(termination.register_from_c(reg))
            _lib.check(L.nsb200_engine_finalize(eng.h, stream))
            _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
        else:
            _lib.check(L.nsb200_engine_init(eng.h, _lib.key_arg(key), ctypes.byref(tc), stream))
            gather = self._gather_tensor(eng) if (world > 1 and not p2p) else None
            lmin = self._contour_agreement(eng) if gather is not None else None
            host_tc = self._effective_host_cond(term_cond) if not plain else None

            def one_body():
                _lib.check(L.nsb200_engine_step_begin(eng.h, stream))
                if gather is not None:
                    # host-issued collective (NSB200_P2P=0, or no peer access): NCCL all-gather of the packed rows
                    import torch.distributed as dist
                    rows = gather.shape[0] // world
                    dist.all_gather_into_tensor(gather, gather[self._rank * rows:(self._rank + 1) * rows])
                    lmin.all_reduce()
                # p2p: the slice kernel already stored the rows into every rank's buffer; step_end starts with the
                # device-side arrival barrier
                _lib.check(L.nsb200_engine_step_end(eng.h, stream))

            if plain:
                # Bodies are no-ops on the device once the register says done, so they are enqueued in bursts
                # between blocking reads of the register.  The register is replicated and deterministic, so every
                # rank sees `done` after the same burst and the all-gathers stay matched across ranks (an
                # asynchronous poll would let ranks launch different numbers of collectives).
                burst = 4
                _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
                dbg = os.environ.get("NSB200_LOOP_DEBUG")
                while not reg.done:
                    t0 = time.perf_counter()
                    for _ in range(burst):
                        one_body()
                    t1 = time.perf_counter()
                    _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
                    if lmin is not None:
                        lmin.check()
                    if dbg and time.perf_counter() - t0 > 0.02:
                        print(f"[loop rank {self._rank}] slow burst at iteration {reg.iteration}: enqueue "
                              f"{1e3 * (t1 - t0):.1f} ms, wait {1e3 * (time.perf_counter() - t1):.1f} ms", file=sys.stderr)
            else:
                _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
                while True:
                    done_now = termination.determine_termination(host_tc, termination.register_from_c(reg))[0] or bool(reg.done)
                    if done_now:
                        break
                    one_body()
                    _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
                    if lmin is not None:
                        lmin.check()
            _lib.check(L.nsb200_engine_finalize(eng.h, stream))
        if p2p:
            err = ctypes.c_int32()
            _lib.check(L.nsb200_engine_p2p_error(eng.h, ctypes.byref(err)))
            if err.value:
                raise RuntimeError("nsb200: a peer GPU did not reach the all-gather barrier (fused NVLink exchange)")
        register = termination.register_from_c(reg)
        if plain:
            termination_reason = int(reg.termination_reason)
        else:
            termination_reason = termination.determine_termination(host_tc, register)[1]
        state = self._state(eng)
        self.last_register = reg
        ms = ctypes.c_double()
        nsl = ctypes.c_int64()
        nall = ctypes.c_int64()
        _lib.check(L.nsb200_engine_slice_profile(eng.h, ctypes.byref(ms), ctypes.byref(nsl), ctypes.byref(nall)))
        self.last_profile = dict(slice_ms=ms.value, slice_launches=nsl.value, all_launches=nall.value,
                                 iterations=int(reg.iteration))
        return termination_reason, register, state

    @staticmethod
    def _print_register(r: TerminationRegister):
        """The per-iteration printout of verbose=True (sharded_static.py:519-549), same fields and wording."""
        ecr, ec = r.evidence_calc_with_remaining, r.evidence_calc
        log_Z_mean, log_Z_var = linear_to_log_stats(ecr.log_Z_mean, log_f2_mean=ecr.log_Z2_mean)
        log_Z_mean0, log_Z_var0 = linear_to_log_stats(ec.log_Z_mean, log_f2_mean=ec.log_Z2_mean)
        with np.errstate(all="ignore"):
            ess = float(np.exp(np.float64(2.0) * ecr.log_Z_mean - ecr.log_dZ2_mean))
            print("-------\n"
                  f"Num samples: {r.num_samples_used}\n"
                  f"Num likelihood evals: {r.num_likelihood_evaluations}\n"
                  f"Efficiency: {r.efficiency}\n"
                  f"log(L) contour: {r.log_L_contour}\n"
                  f"log(Z) est.: {log_Z_mean} +- {float(np.sqrt(np.float64(log_Z_var)))}\n"
                  f"log(Z | remaining) est.: {log_Z_mean - log_Z_mean0} +- "
                  f"{float(np.sqrt(np.float64(log_Z_var) + np.float64(log_Z_var0)))}\n"
                  f"ESS: {ess}\n", flush=True)

    def _connect_peers(self, eng, world) -> bool:
        """Wire the engines of all ranks together for the fused all-gather (include/nsb200.h, nsb200_engine_p2p_*):
        CUDA IPC handles of the engine arenas are exchanged once per engine through the process group; every rank
        must succeed, otherwise all of them keep the host-issued NCCL all-gather (NSB200_P2P=0 forces that path).
        Peer mappings and exported arenas are cached per process by the library (csrc/nsb200.cu open_peer_arena),
        so a sampler rebuilt for every run re-wires in ~1 ms instead of ~15 ms."""
        if getattr(self, "_p2p_state", None) is not None:
            return self._p2p_state
        import torch.distributed as dist
        L = _lib.lib()
        if os.environ.get("NSB200_P2P", "1") != "1" or world > 8 or dist.get_backend() != "nccl":
            self._p2p_state = False  # same decision on every rank (launcher environment), no communication needed
            return False
        ok = True
        handle = (ctypes.c_uint8 * 64)()
        offs = (ctypes.c_int64 * 3)()
        if ok and L.nsb200_engine_p2p_export(eng.h, handle, offs) != 0:
            ok = False
        mine = (bytes(handle), [int(x) for x in offs], bool(ok))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        ok = all(o[2] for o in everyone)
        if ok:
            hbuf = (ctypes.c_uint8 * (64 * world)).from_buffer_copy(b"".join(o[0] for o in everyone))
            obuf = (ctypes.c_int64 * (3 * world))(*[x for o in everyone for x in o[1]])
            ok = L.nsb200_engine_p2p_connect(eng.h, hbuf, obuf) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        _lib.check(L.nsb200_engine_p2p_enabled(eng.h, ctypes.c_int32(1 if ok else 0)))
        self._p2p_state = ok
        return ok

    def _contour_agreement(self, eng):
        """North star: "L_min is agreed each iteration by an NCCL all-reduce".  The sorted live log L is replicated, so
        every rank derives the same contour on its own; with the host-issued collective the ranks still reduce
        (min, max) of it by NCCL once per body and compare with their own value, so replicas that drifted apart stop
        with an error instead of merging different shells.  (The fused path does the same inside k_peer_barrier.)"""
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().nsb200_engine_contour(eng.h, ctypes.byref(ptr)))
        return ContourAgreement(_view(ptr.value, (1,), "<f8", eng), self._rank)

    def _initial_points_external(self, key):
        """create_init_state's prior draws (common/initialisation.py:47-60, common/uniform_sample.py:12-60) with the
        caller's likelihood: round r redraws the rows whose log L is still -inf."""
        from jaxns_b200 import random
        L = _lib.lib()
        N, D = self.num_live_points, self.model.U_ndims
        sample_key = random.split(key, 2)[1]
        d = self.model.desc()
        U = torch.empty((N, D), dtype=torch.float64, device="cuda")
        X = torch.empty((N, D), dtype=torch.float64, device="cuda")
        nev = torch.ones(N, dtype=torch.int64, device="cuda")
        st = _lib.stream_arg()
        _lib.check(L.nsb200_init_propose(ctypes.byref(d), _lib.key_arg(sample_key), ctypes.c_int64(N), ctypes.c_int64(0),
                                         ctypes.c_int64(N), ctypes.c_int32(0), ctypes.c_void_p(0), _lib.ptr(U),
                                         _lib.ptr(X), st))
        logL = self.model.external_log_likelihood(U, X)
        rnd = 0
        while True:
            need = torch.isneginf(logL)  # `while log_L <= -inf` (uniform_sample.py:40-43)
            if not bool(need.any().item()):
                break
            rnd += 1
            if rnd > 10000:
                raise RuntimeError("could not draw initial live points with finite log-likelihood")
            need8 = need.to(torch.uint8).contiguous()
            _lib.check(L.nsb200_init_propose(ctypes.byref(d), _lib.key_arg(sample_key), ctypes.c_int64(N),
                                             ctypes.c_int64(0), ctypes.c_int64(N), ctypes.c_int32(rnd), _lib.ptr(need8),
                                             _lib.ptr(U), _lib.ptr(X), st))
            logL = torch.where(need, self.model.external_log_likelihood(U, X), logL)
            nev += need.to(torch.int64)
        return U, logL.contiguous(), nev

    def _run_external(self, eng, key, tc, reg, stream, world, host_tc):
        """The loop with a caller-evaluated likelihood: every body is step_begin (discard + append), the split
        slice rounds around model.call_likelihood, the all-gather (world > 1) and step_end."""
        L = _lib.lib()
        U0, logL0, nev0 = self._initial_points_external(key)
        _lib.check(L.nsb200_engine_init_external(eng.h, _lib.key_arg(key), ctypes.byref(tc), _lib.ptr(U0), _lib.ptr(logL0),
                                                 _lib.ptr(nev0), stream))
        gather = self._gather_tensor(eng) if world > 1 else None
        lmin = self._contour_agreement(eng) if world > 1 else None
        D = self.model.U_ndims
        n = int(self.num_live_points * self.shell_fraction) // world
        P = self.sampler.split_proposals
        burst = max(4, self.sampler.num_slices // (4 * min(P, 4)))
        # buffers (and the captured graph of a burst, below) live as long as the engine: a second run replays at once
        cache = getattr(self, "_ext_cache", None)
        if cache is None or cache["key"] != (id(eng), P, n, D, burst):
            cache = dict(key=(id(eng), P, n, D, burst), graph=None,
                         prop_U=torch.full((P * n, D), 0.5, dtype=torch.float64, device="cuda"),  # [P, n, D]
                         prop_X=torch.zeros((P * n, D), dtype=torch.float64, device="cuda"),
                         active=torch.zeros(1, dtype=torch.int64, device="cuda"))
            self._ext_cache = cache
        prop_U, prop_X, active = cache["prop_U"], cache["prop_X"], cache["active"]
        grad_pts = torch.empty((n, D), dtype=torch.float64, device="cuda") if self.sampler.gradient_flags else None
        _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))

        def one_burst(st):
            for r in range(burst):
                if grad_pts is not None:  # uni_slice_sampler.py:202-214, :255-269: gradients between the kernels
                    _lib.check(L.nsb200_engine_split_grad_points(eng.h, _lib.ptr(grad_pts), st))
                    grad = self.model.grad_U(grad_pts)
                    _lib.check(L.nsb200_engine_split_grad_begin(eng.h, _lib.ptr(grad), _lib.ptr(prop_U),
                                                                _lib.ptr(prop_X), ctypes.c_void_p(0), st))
                logL = self.model.external_log_likelihood(prop_U, prop_X)
                last = r == burst - 1
                if last:
                    active.zero_()
                _lib.check(L.nsb200_engine_split_accept(eng.h, _lib.ptr(logL), _lib.ptr(prop_U), _lib.ptr(prop_X),
                                                        _lib.ptr(active) if last else ctypes.c_void_p(0), st))

        use_graph = (os.environ.get("NSB200_SPLIT_GRAPH", "1") != "0"
                     and (grad_pts is None or os.environ.get("NSB200_SPLIT_GRAPH_GRAD", "1") != "0")
                     and not getattr(getattr(self.model.log_likelihood, "fn", None), "_nsb200_host_callback", False))
        graph = cache["graph"] if use_graph else None
        try_graph = use_graph and graph is None
        while True:
            if host_tc is None:
                done = bool(reg.done)
            else:
                done = termination.determine_termination(host_tc, termination.register_from_c(reg))[0] or bool(reg.done)
            if done:
                break
            _lib.check(L.nsb200_engine_step_begin(eng.h, stream))
            _lib.check(L.nsb200_engine_split_begin(eng.h, _lib.ptr(prop_U), _lib.ptr(prop_X), stream))
            while True:
                if graph is not None:
                    graph.replay()
                else:
                    one_burst(stream)
                    if try_graph:
                        # the burst is launch-bound (a dozen small kernels per round): capture it once, replay it for
                        # the rest of the run.  Likelihoods that synchronise or call back to the host cannot be
                        # captured; they keep the eager loop.
                        try_graph = False
                        try:
                            graph = cache["graph"] = self._capture(one_burst)
                        except Exception as exc:  # noqa: BLE001
                            graph = None
                            try:
                                torch.cuda.synchronize()
                            except Exception:  # noqa: BLE001
                                pass
                            warnings.warn(f"nsb200: the likelihood rounds could not be captured in a CUDA graph ({exc}); "
                                          "continuing with eager launches (NSB200_SPLIT_GRAPH=0 silences this)")
                if world > 1:  # every rank leaves the rounds together (the all-gather below must stay matched)
                    import torch.distributed as dist
                    dist.all_reduce(active, op=dist.ReduceOp.MAX)
                n_active = int(active.item())
                if n_active & (1 << 62):
                    raise RuntimeError("nsb200: a slice chain did not accept within 65536 proposals: the likelihood is "
                                       "non-deterministic or NaN at its seed point")
                if n_active == 0:
                    break
            _lib.check(L.nsb200_engine_split_finish(eng.h, stream))
            if world > 1:
                import torch.distributed as dist
                rows = gather.shape[0] // world
                dist.all_gather_into_tensor(gather, gather[self._rank * rows:(self._rank + 1) * rows])
                lmin.all_reduce()
            _lib.check(L.nsb200_engine_step_end(eng.h, stream))
            _lib.check(L.nsb200_engine_register(eng.h, ctypes.byref(reg), stream))
            if lmin is not None:
                lmin.check()
        _lib.check(L.nsb200_engine_finalize(eng.h, stream))

    @staticmethod
    def _capture(fn):
        """Capture fn(stream) into a CUDA graph on a side stream (torch.cuda.graph() without its gc / cache flush)."""
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g.capture_begin()
            try:
                fn(_lib.stream_arg())
            finally:
                g.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        return g

    def _effective_host_cond(self, term_cond):
        """max_samples lowered by one iteration's space (sharded_static.py:464-470), applied to every leaf."""
        m = int(self.num_live_points * self.shell_fraction)
        lim = self.max_samples - m * (1 + self.sampler.num_phantom())

        def fix(c):
            if isinstance(c, TerminationCondition):
                if c.max_samples is not None:
                    return c._replace(max_samples=min(float(termination._f(c.max_samples)), lim))
                return c
            return type(c)(conds=[fix(x) for x in c.conds])

        return fix(term_cond)

    def _gather_tensor(self, eng) -> torch.Tensor:
        buf = ctypes.c_void_p()
        rows = ctypes.c_int64()
        rd = ctypes.c_int64()
        _lib.check(_lib.lib().nsb200_engine_gather_buffer(eng.h, ctypes.byref(buf), ctypes.byref(rows), ctypes.byref(rd)))
        world = len(self.devices)
        return _view(buf.value, (rows.value * world, rd.value), "<f8", eng)

    def _state(self, eng) -> NestedSamplerState:
        v = _lib.NsStateView()
        _lib.check(_lib.lib().nsb200_engine_state(eng.h, ctypes.byref(v), _lib.stream_arg()))
        cap, D = v.capacity, v.D
        # The state owns its arrays (device-to-device copies, ~30 us for the 92 MB store of config 2): the engine's
        # arena is refilled by the next run of the same sampler, and the reference returns immutable arrays -- a state
        # kept from an earlier run must stay valid (tests/test_gpu_parity.py::test_state_survives_the_next_run).
        sc = SampleCollection(
            sender_node_idx=_view(v.sender_node_idx, (cap,), "<i8", eng).clone(),
            log_L=_view(v.log_L, (cap,), "<f8", eng).clone(),
            U_samples=_view(v.U_samples, (cap, D), "<f8", eng).clone(),
            num_likelihood_evaluations=_view(v.num_likelihood_evaluations, (cap,), "<i8", eng).clone(),
            phantom=_view(v.phantom, (cap,), "|u1", eng).bool(),
        )
        return NestedSamplerState(key=np.array([v.key[0], v.key[1]], dtype=np.uint32),
                                  next_sample_idx=int(v.next_sample_idx), num_samples=int(v.num_samples),
                                  sample_collection=sc)

    # ------------------------------------------------------------------------------------------
    def _to_results(self, termination_reason, state: NestedSamplerState, trim: bool) -> NestedSamplerResults:
        """_to_results (sharded_static.py:652-773)."""
        sc = state.sample_collection
        capacity = sc.log_L.numel()
        num_samples = min(int(state.num_samples), capacity)
        if trim:
            sc = SampleCollection(*[x[:num_samples] for x in sc])
            counts = count_crossed_edges(SampleTreeGraph(sc.sender_node_idx, sc.log_L))
        else:
            counts = count_crossed_edges(SampleTreeGraph(sc.sender_node_idx, sc.log_L), num_samples=num_samples)
        idx = counts.samples_indices
        num_live_points = counts.num_live_points
        log_L = sc.log_L[idx]
        U_samples = sc.U_samples[idx]
        num_likelihood_evaluations = sc.num_likelihood_evaluations[idx]
        final, per = compute_evidence_stats(log_L, num_live_points, num_samples=None if trim else num_samples)
        log_Z_mean, log_Z_var = linear_to_log_stats(final.log_Z_mean, log_f2_mean=final.log_Z2_mean)
        log_Z_uncert = math.sqrt(log_Z_var)
        total_phantom = int(sc.phantom.sum().item())
        phantom_fraction = total_phantom / num_samples
        k = phantom_fraction / (1. - phantom_fraction)
        log_Z_uncert = log_Z_uncert * math.sqrt(1. + k)
        ESS = effective_sample_size_kish(final.log_Z_mean, final.log_dZ2_mean) / (1. + k)
        samples = self.model.transform(U_samples)
        # dp = normalise_log_space(LogSpace(log_dZ_mean)) (internals/log_semiring.py:359-378)
        norm = logsumexp(per.log_dZ_mean)
        log_dp = per.log_dZ_mean - norm
        if norm == -math.inf:
            log_dp = torch.full_like(log_dp, -math.inf)
        # H estimators (:728-737); sums are small reductions over M handled by torch on the device
        ll = torch.where(torch.isneginf(log_dp), torch.zeros_like(log_L), log_L)
        H_instable = -(float((torch.exp(log_dp) * ll).sum().item()) - log_Z_mean)
        H_stable = -float((torch.exp(log_dp) * (-per.log_X_mean)).sum().item())
        H_mean = H_instable if math.isfinite(H_instable) else H_stable
        total_evals = int(num_likelihood_evaluations.sum().item())
        log_eff = math.log(num_samples) - (math.log(total_evals) if total_evals > 0 else -math.inf)
        log_post = log_L + self.model.log_prob_prior(U_samples)
        return NestedSamplerResults(
            log_Z_mean=log_Z_mean, log_Z_uncert=log_Z_uncert, ESS=ESS, H_mean=H_mean, samples=samples,
            parametrised_samples=self.model.transform_parametrised(U_samples), U_samples=U_samples, log_L_samples=log_L, log_dp_mean=log_dp,
            log_X_mean=per.log_X_mean, log_posterior_density=log_post, num_live_points_per_sample=num_live_points,
            num_likelihood_evaluations_per_sample=num_likelihood_evaluations, total_num_samples=num_samples,
            total_phantom_samples=total_phantom, total_num_likelihood_evaluations=total_evals,
            log_efficiency=log_eff, termination_reason=termination_reason)

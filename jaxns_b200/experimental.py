"""
Callers of the hot path: global optimisation on top of the static nested-sampling loop, name for
name with /root/reference/src/jaxns/experimental/global_optimisation.py:28-182 and experimental/public.py:20-140.

SimpleGlobalOptimisation IS the nested-sampling engine with three twists the engine supports natively: a dead-point
store of only 10 x num_search_chains rows that wraps around (max_samples=None turns the sample cap off,
sharded_static.py:76-78, global_optimisation.py:163-172), termination on likelihood evaluations / contour / spread /
efficiency, and "the result" = the best stored point instead of an evidence.

gradient_slice chains (uni_slice_sampler.py:202-214, the reference's default for GlobalOptimisation) run through the
split propose / accept kernels with Model.grad_U between them; the fine-tune (global_optimisation.py:56-73) is a
Newton-CG descent of -log L over the quick_unit-unconstrained cube (internals/constraint_bijections.py:11-40), written
here with torch autograd Hessian-vector products -- the reference's own newton_cg_solver is not restated step for step,
so the fine-tuned point agrees with the reference's only to the optimiser's tolerance.

EvidenceMaximisation (experimental/evidence_maximisation.py:40-300): the E-step IS a nested-sampling run of the model at
the current parameters (`get_parameter`, `Prior(...).parametrised()`: jaxns_b200/context.py, framework.py); the M-step
maximises log sum_i exp(log L(U_i; params) + log w_i) over the run's samples with the same Newton-CG.
"""
import dataclasses
import io
import math
from typing import Any, NamedTuple, Optional, TextIO, Union

import torch

from jaxns_b200.constraint_bijections import quick_unit, quick_unit_inverse
from jaxns_b200.nested_sampler import ShardedStaticNestedSampler
from jaxns_b200.samplers import AbstractSampler, UniDimSliceSampler
from jaxns_b200.types import SampleCollection, TerminationCondition

__all__ = ["GlobalOptimisationResults", "GlobalOptimisationTerminationCondition", "GlobalOptimisationState",
           "SimpleGlobalOptimisation", "GlobalOptimisation", "DefaultGlobalOptimisation", "go_summary",
           "gradient_based_optimisation", "quick_unit", "quick_unit_inverse", "newton_cg", "EvidenceMaximisation"]


class GlobalOptimisationState(NamedTuple):
    key: Any
    samples: SampleCollection
    num_samples: int
    relative_spread: float
    absolute_spread: float
    num_likelihood_evaluations: int


class GlobalOptimisationResults(NamedTuple):
    U_solution: Any
    X_solution: Any
    solution: Any
    log_L_solution: float
    log_L_progress: Any
    num_likelihood_evaluations: int
    num_samples: int
    termination_reason: int
    relative_spread: float
    absolute_spread: float


class GlobalOptimisationTerminationCondition(NamedTuple):
    max_likelihood_evaluations: Optional[float] = None
    log_likelihood_contour: Optional[float] = None
    rtol: Optional[float] = None
    atol: Optional[float] = None
    min_efficiency: Optional[float] = None


def newton_cg(loss_of, z0, max_iters: int = 100, cg_iters: int = 50, gtol: float = 1e-10):
    """Minimise a scalar torch function of one tensor by Newton-CG (the role of the reference's newton_cg_solver in
    global_optimisation.py:62 and evidence_maximisation.py:181).  Each outer step solves H p = -g by conjugate
    gradients on autograd Hessian-vector products (stopping at negative curvature or the Eisenstat-Walker residual),
    then backtracks on the step length until the loss decreases.  Returns (z, f(z), CG iterations)."""
    z = z0.detach().clone()
    n_cg = 0
    with torch.no_grad():
        f = float(loss_of(z))
    for _ in range(max_iters):
        zz = z.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            fz = loss_of(zz)
            (g,) = torch.autograd.grad(fz, zz, create_graph=True)
        gn = float(g.detach().norm())
        if not math.isfinite(gn) or gn < gtol:
            break

        def hvp(v):
            (hv,) = torch.autograd.grad(g, zz, grad_outputs=v, retain_graph=True)
            return hv

        p = torch.zeros_like(z)
        r = g.detach().clone()
        d = -r
        rr = float((r * r).sum())
        tol = min(0.5, math.sqrt(gn)) * gn
        for _ in range(cg_iters):
            n_cg += 1
            Hd = hvp(d)
            curv = float((d * Hd).sum())
            if curv <= 0:
                if float(p.abs().sum()) == 0.0:
                    p = -g.detach()  # steepest descent when the first direction already has negative curvature
                break
            alpha = rr / curv
            p = p + alpha * d
            r = r + alpha * Hd
            rr_new = float((r * r).sum())
            if math.sqrt(rr_new) < tol:
                break
            d = -r + (rr_new / rr) * d
            rr = rr_new
        step, improved = 1.0, False
        for _ in range(30):
            with torch.no_grad():
                f_new = float(loss_of(z + step * p))
            if math.isfinite(f_new) and f_new < f:
                z, f, improved = (z + step * p).detach(), f_new, True
                break
            step *= 0.5
        if not improved:
            break
    return z, f, n_cg


def gradient_based_optimisation(model, init_U_point):
    """global_optimisation.py:56-73: minimise -log L(quick_unit(z)) from z0 = quick_unit_inverse(U).  Returns
    (U, log L, number of function evaluations counted as the reference does: 4 per CG iteration)."""
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    z0 = quick_unit_inverse(torch.as_tensor(init_U_point, dtype=torch.float64, device=dev).reshape(1, -1))
    z, f, n_cg = newton_cg(lambda zz: -model.log_likelihood_torch(quick_unit(zz))[0], z0)
    return quick_unit(z)[0].detach(), -f, 4 * n_cg


@dataclasses.dataclass(eq=False)
class SimpleGlobalOptimisation:
    """global_optimisation.py:76-182."""
    sampler: AbstractSampler
    num_search_chains: int
    model: Any
    shell_frac: float = 0.5
    devices: Optional[Any] = None
    verbose: bool = False

    def __post_init__(self):
        if self.num_search_chains < 1:
            raise ValueError("num_search_chains must be >= 1.")
        self.num_search_chains = int(self.num_search_chains)
        self._nested_sampler = ShardedStaticNestedSampler(
            model=self.model,
            max_samples=self.num_search_chains * 10,
            init_efficiency_threshold=0.1,
            sampler=self.sampler,
            num_live_points=self.num_search_chains,
            shell_fraction=self.shell_frac,
            devices=self.devices,
            verbose=self.verbose
        )

    def _gradient_descent(self, results: GlobalOptimisationResults) -> GlobalOptimisationResults:
        """global_optimisation.py:103-115."""
        U_solution, log_L_solution, n_evals = gradient_based_optimisation(self.model, results.U_solution)
        return results._replace(
            U_solution=U_solution, log_L_solution=log_L_solution, X_solution=self.model.transform(U_solution),
            solution=self.model.prepare_input(U_solution),
            num_likelihood_evaluations=results.num_likelihood_evaluations + n_evals)

    def _to_results(self, termination_reason, state: GlobalOptimisationState) -> GlobalOptimisationResults:
        """global_optimisation.py:118-147: best stored point; the store may have wrapped, so only the rows written
        so far (num_samples, capped at the capacity) are looked at."""
        log_L = state.samples.log_L
        cap = log_L.numel()
        is_sample = torch.arange(cap, device=log_L.device) < state.num_samples
        masked = torch.where(is_sample, log_L, torch.full_like(log_L, math.nan))
        best_idx = int(torch.argmax(torch.nan_to_num(masked, nan=-math.inf)).item())
        U_solution = state.samples.U_samples[best_idx]
        X_solution = self.model.transform(U_solution)
        solution = self.model.prepare_input(U_solution)
        log_L_progress = torch.sort(masked).values  # low to high likelihoods, NaN (unused rows) at the end
        return GlobalOptimisationResults(
            U_solution=U_solution, X_solution=X_solution, solution=solution,
            log_L_solution=float(log_L[best_idx].item()), log_L_progress=log_L_progress,
            num_likelihood_evaluations=int(state.num_likelihood_evaluations), num_samples=int(state.num_samples),
            relative_spread=float(state.relative_spread), absolute_spread=float(state.absolute_spread),
            termination_reason=int(termination_reason))

    def _run(self, key, term_cond: GlobalOptimisationTerminationCondition):
        """global_optimisation.py:149-182."""
        termination_reason, termination_register, state = self._nested_sampler._run(
            key=key,
            term_cond=TerminationCondition(
                max_num_likelihood_evaluations=term_cond.max_likelihood_evaluations,
                log_L_contour=term_cond.log_likelihood_contour,
                efficiency_threshold=term_cond.min_efficiency,
                atol=term_cond.atol,
                rtol=term_cond.rtol,
                max_samples=None  # turn off max samples for global optimisation: the store index wraps
            )
        )
        go_state = GlobalOptimisationState(
            key=state.key, samples=state.sample_collection, num_samples=state.num_samples,
            absolute_spread=termination_register.absolute_spread, relative_spread=termination_register.relative_spread,
            num_likelihood_evaluations=termination_register.num_likelihood_evaluations)
        return termination_reason, go_state


def _bit_mask(int_mask, width=8):
    return list(map(int, '{:0{size}b}'.format(int_mask, size=width)))[::-1]


def go_summary(results: GlobalOptimisationResults, f_obj: Optional[Union[str, TextIO]] = None) -> str:
    """global_optimisation.py:200-300: the text report of a global-optimisation run, same lines and rounding."""
    main_s = []

    def _print(msg):
        print(msg)
        main_s.append(msg)

    def _round(v, uncert_v):
        v, uncert_v = float(v), float(uncert_v)
        try:
            sig_figs = -int("{:e}".format(uncert_v).split('e')[1]) + 1
            return round(v, sig_figs)
        except Exception:
            return v

    _print("--------")
    _print("Termination Conditions:")
    names = ['Reached max samples', 'Evidence uncertainty low enough', 'Small remaining evidence', 'Reached ESS',
             "Used max num likelihood evaluations", 'Likelihood contour reached', 'Sampler efficiency too low',
             'All live-points are on a single plateau (sign of possible precision error)',
             'relative spread of live points < rtol', 'absolute spread of live points < atol',
             'no seed points left (consider decreasing shell_fraction)']
    for bit, description in zip(_bit_mask(int(results.termination_reason), width=11), names):
        if bit == 1:
            _print(description)
    _print("--------")
    _print(f"likelihood evals: {int(results.num_likelihood_evaluations):d}")
    _print(f"samples: {int(results.num_samples):d}")
    _print(f"likelihood evals / sample: {float(results.num_likelihood_evaluations / results.num_samples):.1f}")
    _print("--------")
    _print(f"max(log_L)={_round(results.log_L_solution, results.log_L_solution)}")
    _print(f"relative spread: {_round(results.relative_spread, results.relative_spread)}")
    _print(f"absolute spread: {_round(results.absolute_spread, results.absolute_spread)}")
    for name, value in results.X_solution.items():
        v = value.detach().cpu().numpy()
        if v.size == 0:
            continue
        _print("--------")
        is_shaped = v.ndim > 0
        var_name = f"{name}[{','.join(['#'] * v.ndim)}]" if is_shaped else name
        _print(f"{var_name}: max(L) est.")
        if is_shaped:
            import numpy as np
            for inds in np.indices(v.shape).reshape((v.ndim, -1)).T:
                point = v[tuple(inds)]
                _print(f"{name}[{','.join(str(i) for i in inds)}]: {_round(point, 0.1 * point)}")
        else:
            _print(f"{name}: {_round(v, 0.1 * v)}")
    _print("--------")
    out = "\n".join(main_s)
    if f_obj is not None:
        if isinstance(f_obj, str):
            with open(f_obj, "w") as f:
                f.write(out)
        elif isinstance(f_obj, io.TextIOBase):
            f_obj.write(out)
        else:
            raise TypeError(f"Invalid f_obj: {type(f_obj)}")
    return out


@dataclasses.dataclass(eq=False)
class GlobalOptimisation:
    """experimental/public.py:20-140, defaults included (gradient_slice=True: 15 D chains, s = 2)."""
    model: Any
    num_search_chains: Optional[int] = None
    s: Optional[int] = None
    k: Optional[int] = None
    gradient_slice: bool = True
    shell_frac: Optional[float] = None
    devices: Optional[Any] = None
    verbose: bool = False

    def __post_init__(self):
        if self.num_search_chains is None:
            self.num_search_chains = self.model.U_ndims * (15 if self.gradient_slice else 100)
        if self.s is None:
            self.s = 2 if self.gradient_slice else 10
        if self.shell_frac is None:
            self.shell_frac = 0.5
        if self.k is None:
            self.k = self.model.U_ndims * self.s - 1
        sampler = UniDimSliceSampler(model=self.model, num_slices=self.model.U_ndims * int(self.s),
                                     num_phantom_save=int(self.k), midpoint_shrink=True, perfect=True,
                                     gradient_slice=self.gradient_slice)
        self._global_optimiser = SimpleGlobalOptimisation(
            sampler=sampler, num_search_chains=int(self.num_search_chains), shell_frac=float(self.shell_frac),
            model=self.model, devices=self.devices, verbose=self.verbose)
        self.summary = go_summary

    def __call__(self, key, term_cond: Optional[GlobalOptimisationTerminationCondition] = None,
                 finetune: bool = False) -> GlobalOptimisationResults:
        if term_cond is None:
            term_cond = GlobalOptimisationTerminationCondition(min_efficiency=3e-2)
        termination_reason, state = self._global_optimiser._run(key, term_cond)
        results = self._global_optimiser._to_results(termination_reason, state)
        if finetune:
            results = self._global_optimiser._gradient_descent(results=results)
        return results


DefaultGlobalOptimisation = GlobalOptimisation


class MStepData(NamedTuple):
    U_samples: Any
    log_weights: Any


@dataclasses.dataclass(eq=False)
class EvidenceMaximisation:
    """experimental/evidence_maximisation.py:40-300, same fields, defaults, convergence rules and printed lines.

    E-step: NestedSampler(model(params), **ns_kwargs) run to `termination_cond`, trimmed results (:84-112) -- the hot
    path.  M-step (:149-221): with the run's samples U_i and weights log w_i = log dp_i - log L_i + log Z, maximise
    log Z(params) = logsumexp_i(log L(U_i; params) + log w_i) by Newton-CG, one solve per epoch, until no parameter
    moved by more than `gtol` (or max_num_epochs).  The reference pads the samples to a power of two with zero-weight
    rows so that XLA re-compiles less often; there is no compilation here, so nothing is padded."""
    model: Any
    ns_kwargs: Optional[dict] = None
    max_num_epochs: int = 50
    gtol: float = 1e-2
    log_Z_ftol: float = 1.
    log_Z_atol: float = 1e-4
    batch_size: Optional[int] = 128
    termination_cond: Optional[TerminationCondition] = None
    verbose: bool = False

    def __post_init__(self):
        if self.ns_kwargs is None:
            self.ns_kwargs = {}

    def e_step(self, key, params, desc: str):
        """The E-step is just nested sampling (:114-128)."""
        from jaxns_b200.public import NestedSampler
        print(f"Running E-step... {desc}")
        ns = NestedSampler(model=self.model(params={k: v.detach() for k, v in params.items()}), **self.ns_kwargs)
        termination_reason, state = ns(key, self.termination_cond)
        return ns.to_results(termination_reason=termination_reason, state=state, trim=True)

    def _log_evidence(self, params, data: MStepData):
        log_dZ = self.model(params=params).log_likelihood_torch(data.U_samples) + data.log_weights
        return torch.logsumexp(log_dZ, dim=0)

    def _m_step(self, key, params, data: MStepData):
        """One Newton-CG solve of -log Z over the flattened parameters (:174-188)."""
        names = list(params.keys())
        shapes = [params[k].shape for k in names]
        sizes = [int(params[k].numel()) for k in names]
        flat0 = torch.cat([params[k].detach().reshape(-1) for k in names]) if names else torch.zeros(0, dtype=torch.float64)

        def unflatten(z):
            out, o = {}, 0
            for k, shp, sz in zip(names, shapes, sizes):
                out[k] = z[o:o + sz].reshape(shp)
                o += sz
            return out

        def loss(z):
            log_Z = self._log_evidence(unflatten(z), data)
            if self.verbose:
                print(f"log_Z={float(log_Z)}")
            return -log_Z

        if flat0.numel() == 0:
            return params, float(-loss(flat0))
        z, f, _ = newton_cg(loss, flat0)
        return {k: v.detach() for k, v in unflatten(z).items()}, -f

    def m_step(self, key, params, ns_results, desc: str):
        """:190-233."""
        num_samples = int(ns_results.total_num_samples)
        print(f"Running M-step ({num_samples} samples)... {desc}")
        log_weights = ns_results.log_dp_mean - ns_results.log_L_samples + ns_results.log_Z_mean
        data = MStepData(U_samples=ns_results.U_samples, log_weights=log_weights)
        last_params = params
        epoch = 0
        log_Z = None
        while epoch < self.max_num_epochs:
            params, log_Z = self._m_step(key=key, params=params, data=data)
            l_oo = {k: (float((params[k] - last_params[k]).abs().max()) if params[k].numel() > 0 else 0.) for k in params}
            last_params = params
            print(f"{desc}: Epoch {epoch}: log_Z={log_Z}, l_oo={l_oo}")
            if all(v < self.gtol for v in l_oo.values()):
                break
            epoch += 1
        return params, log_Z

    def train(self, num_steps: int = 10, params=None):
        """:235-300: alternate E and M steps until log Z stops changing (atol, or ftol x its uncertainty)."""
        from jaxns_b200 import random
        if params is None:
            params = self.model.params
        log_Z = -math.inf
        ns_results = None
        for step in range(num_steps):
            key_e_step, key_m_step = random.split(random.PRNGKey(step), 2)
            if ns_results is None:
                desc = f"Step {step}: Initial run"
            else:
                desc = f"Step {step}: log Z = {float(ns_results.log_Z_mean):.4f} +- {float(ns_results.log_Z_uncert):.4f}"
            ns_results = self.e_step(key=key_e_step, params=params, desc=desc)
            log_Z_change = abs(float(ns_results.log_Z_mean) - log_Z)
            if log_Z_change < self.log_Z_atol:
                break
            if log_Z_change < float(self.log_Z_ftol * float(ns_results.log_Z_uncert)):
                break
            log_Z = float(ns_results.log_Z_mean)
            desc = f"Step {step}: log Z = {float(ns_results.log_Z_mean):.4f} +- {float(ns_results.log_Z_uncert):.4f}"
            params, log_Z_opt = self.m_step(key=key_m_step, params=params, ns_results=ns_results, desc=desc)
        if ns_results is None:
            raise RuntimeError("No results were computed.")
        return ns_results, params

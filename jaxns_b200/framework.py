"""
Model / Prior with the reference's signatures (/root/reference/src/jaxns/framework/model.py:24-208,
framework/prior.py:65-161, framework/ops.py:21-36,240-326) for registered likelihood families.

`prior_model` stays a generator function that yields Prior objects and returns the likelihood
inputs.  `log_likelihood` is either a RegisteredLikelihood (jaxns_b200.likelihoods: fused into the
slice kernel) or any batched device callable (wrapped in ExternalLikelihood): then the slice step is
split into propose / accept kernels around the call, still entirely on the GPU.
"""
import ctypes
from typing import Callable, List, Optional

import numpy as np
import torch

from jaxns_b200 import _lib, context, distributions
from jaxns_b200.constraint_bijections import quick_unit
from jaxns_b200.likelihoods import ExternalLikelihood, RegisteredLikelihood

__all__ = ["Prior", "SingularPrior", "Model"]


class _NeedsGeneral(Exception):
    """The prior model does arithmetic on its variables or feeds them into other priors: it cannot be reduced to the
    static per-dimension quantile arrays and is evaluated as batched torch code instead."""


class _Var:
    """Placeholder sent into the prior_model generator for a yielded Prior (static analysis pass)."""
    _nsb200_placeholder = True

    def __init__(self, index: int, size: int, name: Optional[str]):
        self.index, self.size, self.name = index, size, name

    def _general(self, *args, **kwargs):
        raise _NeedsGeneral()

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _general
    __pow__ = __rpow__ = __neg__ = __abs__ = __getitem__ = __matmul__ = __rmatmul__ = __array__ = _general
    __iter__ = __len__ = __float__ = _general


class Prior:
    """Prior(dist_or_value, name=None) (framework/prior.py:65-161)."""

    singular = False

    def __init__(self, dist_or_value, name: Optional[str] = None):
        self.name = name
        self.dist = distributions.from_any(dist_or_value)

    def parametrised(self, random_init: bool = False) -> "SingularPrior":
        """framework/prior.py:146-199: the prior becomes a point value driven by a free parameter `<name>_param` on
        the real line (median of the prior at 0), mapped through quick_unit and the prior's quantile; it keeps the
        prior's log density at that value.  Must be called inside prior_model (it asks the context for the parameter)."""
        import warnings
        if self.name is None:
            raise ValueError("Prior must have a name to be parametrised.")
        size = self.dist.event_size()
        if size == 0:
            warnings.warn(f"Creating a zero-sized parameter for {self.name}. Probably unintended.")

        def init(shape, dtype):
            if random_init:
                from jaxns_b200 import random
                return random.normal(context.next_rng_key(), int(np.prod(shape)))
            return np.zeros(shape)

        param = context.get_parameter(f"{self.name}_param", (size,), np.float64, init=init)
        value = self.dist.quantile_torch(quick_unit(param).reshape(1, -1))
        return SingularPrior(value=value, base_prior=self, name=self.name)


class SingularPrior(Prior):
    """framework/prior.py:32-62: no U dimensions, the value itself, the base prior's log density at it."""
    singular = True

    def __init__(self, value, base_prior: Prior, name: Optional[str] = None):
        self.name = name
        self.value = value
        self.base_prior = base_prior
        self.dist = distributions.Constant(value)

    def __repr__(self):
        return f"{self.value} -> {self.base_prior}"


class Model:
    """Model(prior_model, log_likelihood, params=None) (framework/model.py:29-41)."""

    def __init__(self, prior_model: Callable, log_likelihood, params=None):
        if not isinstance(log_likelihood, RegisteredLikelihood):
            log_likelihood = ExternalLikelihood(log_likelihood)
        self.prior_model = prior_model
        self.log_likelihood = log_likelihood
        self.is_external = isinstance(log_likelihood, ExternalLikelihood)
        self.is_general = False
        self._priors: List[Prior] = []
        self._params = {}
        self._dev = None
        if params is not None and not self.is_external:
            raise NotImplementedError("Parametrised models need a callable likelihood.")
        if self.is_external:
            # framework/model.py:33-35: without params the model initialises them (one pass through the prior model
            # and the likelihood in an initialising context, as parse_joint does)
            dev = "cuda" if torch.cuda.is_available() else "cpu"
            self._params = ({k: torch.as_tensor(v, dtype=torch.float64, device=dev) for k, v in params.items()}
                            if params is not None else self.init_params())
            if self._params:
                self._analyse_general()
                return
        try:
            with context.bind({}):
                self._analyse_static()
        except _NeedsGeneral:
            if not self.is_external:
                raise NotImplementedError(
                    "Registered likelihood families are fused with per-dimension Uniform / Normal quantile transforms; "
                    "dependent, mixed or dense-MVN priors need a callable likelihood (evaluated between the propose and "
                    "accept kernels).")
            self._analyse_general()

    def init_params(self, rng=None):
        """framework/model.py:93-115: every get_parameter the prior model and the likelihood ask for, at its `init`."""
        with context.bind(None, rng) as ctx:
            ret, _, _, _ = self._run_generator(None, bound=False)
            self.log_likelihood.fn(*ret)
        return {k: v.detach() for k, v in ctx.params.items()}

    def set_params(self, params) -> "Model":
        """framework/model.py:62-73: a new model with these parameter values."""
        return Model(prior_model=self.prior_model, log_likelihood=self.log_likelihood, params=params)

    def __call__(self, params) -> "Model":
        return self.set_params(params=params)

    def __repr__(self):
        return f"Model(U_ndims={self.U_ndims}, num_params={self.num_params})"

    def _analyse_static(self):
        """Drive the generator once with placeholders: works when every prior has constant parameters and the model
        returns its variables as they are -- then U -> X is the kernels' per-dimension quantile transform."""
        self._priors = []
        gen = self.prior_model()
        try:
            p = next(gen)
            while True:
                if not isinstance(p, Prior):
                    raise TypeError(f"prior_model must yield Prior objects, got {type(p)}")
                if p.dist.dynamic:
                    raise _NeedsGeneral()
                v = _Var(len(self._priors), p.dist.event_size(), p.name)
                self._priors.append(p)
                p = gen.send(v)
        except StopIteration as stop:
            ret = stop.value
        ret = ret if isinstance(ret, tuple) else (ret,)
        if not all(isinstance(r, _Var) for r in ret):
            if self.is_external:
                raise _NeedsGeneral()
            raise NotImplementedError("prior_model must return (a tuple of) its yielded variables.")
        if not self.is_external and [r.index for r in ret] != list(range(len(self._priors))):
            raise NotImplementedError("prior_model must return its yielded variables in order: the registered "
                                      "likelihood consumes their concatenation.")
        offs = np.concatenate([[0], np.cumsum([p.dist.event_size() for p in self._priors])]).astype(int)
        self._ret_slices = [(int(offs[r.index]), int(offs[r.index + 1])) for r in ret]
        kinds = {p.dist.prior_kind for p in self._priors}
        if len(kinds) != 1:
            if self.is_external:
                raise _NeedsGeneral()
            raise NotImplementedError("Mixing Uniform and Normal priors in one model needs a callable likelihood.")
        self._prior_kind = kinds.pop()
        self._a = np.concatenate([p.dist.quantile_params()[1] for p in self._priors])
        self._b = np.concatenate([p.dist.quantile_params()[2] for p in self._priors])
        self._D = int(self._a.size)
        self._params_host = np.ascontiguousarray(self.log_likelihood.pack(self._D), np.float64)

    def _analyse_general(self):
        """General models (framework/ops.py:240-326 with arbitrary generators): the prior transform IS the generator,
        run on batched device tensors -- every yielded Prior's quantile at its slice of U, its value sent back in, the
        returned expressions handed to the likelihood.  The slice kernels only ever see U (propose / accept), so
        dependent priors (tests/conftest.py:144-177 `basic3`), mixed families, dense MVN priors, constants and
        re-ordered or derived return values all work; the kernels get a Uniform(0, 1) descriptor of the right width."""
        self.is_general = True
        _, named, _, sizes = self._run_generator(None)
        self._D = int(sum(sizes))
        self._names = list(named.keys())
        self._prior_kind = distributions.Uniform.prior_kind
        self._a = np.zeros(self._D)
        self._b = np.ones(self._D)
        self._params_host = np.zeros(0)
        self._ret_slices = None

    def _run_generator(self, U, bound: bool = True, parametrised: Optional[dict] = None):
        """One batched pass through the prior model.  U [n, D] on the device (None: a shape-discovery pass at U = 1/2).
        Returns (likelihood inputs, {name: X}, log prior density [n], event sizes).  `bound`: open a context with the
        model's parameters (False: the caller already holds one); `parametrised` collects the singular priors' values."""
        if bound:
            with context.bind(self._params):
                return self._run_generator(U, False, parametrised)
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        probe = U is None
        n = 1 if probe else U.shape[0]
        gen = self.prior_model()
        named, sizes, o = {}, [], 0
        log_prior = torch.zeros(n, dtype=torch.float64, device=dev if probe else U.device)
        try:
            p = next(gen)
            while True:
                if not isinstance(p, Prior):
                    raise TypeError(f"prior_model must yield Prior objects, got {type(p)}")
                size = p.dist.event_size()
                u = torch.full((1, size), 0.5, dtype=torch.float64, device=dev) if probe else U[:, o:o + size]
                x = p.dist.quantile_torch(u)
                if x.shape[0] != n:
                    x = x.expand(n, x.shape[-1])
                if p.singular:
                    log_prior = log_prior + p.base_prior.dist.log_prob_torch(x)
                    if parametrised is not None and p.name is not None:
                        parametrised[p.name] = x
                else:
                    log_prior = log_prior + p.dist.log_prob_torch(x)
                    if p.name is not None:
                        named[p.name] = x
                sizes.append(size)
                o += size
                p = gen.send(x)
        except StopIteration as stop:
            ret = stop.value
        ret = ret if isinstance(ret, tuple) else (ret,)
        return ret, named, log_prior, sizes

    def _general_log_likelihood(self, U: torch.Tensor) -> torch.Tensor:
        with context.bind(self._params), torch.no_grad():
            ret, _, _, _ = self._run_generator(U, bound=False)
            out = self.log_likelihood.fn(*ret)
        out = torch.as_tensor(out, dtype=torch.float64, device=U.device)
        out = out.expand(U.shape[0]) if out.dim() == 0 else out.reshape(-1)
        if out.numel() != U.shape[0]:
            raise ValueError(f"log_likelihood must return one value per row: got {out.numel()} for {U.shape[0]} rows")
        return torch.nan_to_num(out, nan=-float("inf"), posinf=float("inf"), neginf=-float("inf")).contiguous()

    def external_log_likelihood(self, U: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
        """log L of a batch of proposals for the split slice step: U [n, D] in the unit cube, X = the kernels'
        per-dimension transform of U (what static models consume; general models run their own transform on U)."""
        if self.is_general:
            return self._general_log_likelihood(U)
        if not self.is_external:  # a registered family driven through the split step (gradient variants): k_forward
            return self._forward_batch(U.contiguous(), False)[0]
        return self.call_likelihood(X)

    def grad_U(self, U: torch.Tensor) -> torch.Tensor:
        """d log L / dU at a batch U [n, D]: jax.grad(model.forward) of the reference (uni_slice_sampler.py:135) as
        one reverse pass over the whole batch (rows are independent, so the gradient of the summed log L is the
        per-row gradient).  The prior transform and the likelihood are replayed as differentiable torch code; the
        likelihood VALUES the chains are accepted on still come from the kernels / the caller's function."""
        U = U.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            (g,) = torch.autograd.grad(self.log_likelihood_torch(U).sum(), U)
        return g.contiguous()

    def log_likelihood_torch(self, U: torch.Tensor) -> torch.Tensor:
        """log L [n] at U [n, D] as differentiable torch code (prior transform included)."""
        if self.is_general:
            with context.bind(self._params):
                ret, _, _, _ = self._run_generator(U, bound=False)
                out = self.log_likelihood.fn(*ret)
        else:
            a = distributions._t(self._a, U)
            b = distributions._t(self._b, U)
            q = U if self._prior_kind == distributions.Uniform.prior_kind else torch.special.ndtri(U)
            X = a + b * q
            if self.is_external:
                out = self.log_likelihood.fn(*[X[:, lo:hi] for lo, hi in self._ret_slices])
            else:
                out = self.log_likelihood.log_prob_torch(X)
        wants_grad = U.requires_grad or any(p.requires_grad for p in self._params.values())
        if not isinstance(out, torch.Tensor) or (wants_grad and not out.requires_grad):
            raise TypeError("gradient_slice / gradient_guided / finetune need a log_likelihood made of differentiable "
                            "torch operations (a jaxify_likelihood host function has no gradient)")
        out = out.to(torch.float64)
        return out.expand(U.shape[0]) if out.dim() == 0 else out.reshape(-1)

    # -- reference properties -----------------------------------------------------------------
    @property
    def U_ndims(self) -> int:
        return self._D

    @property
    def U_placeholder(self):
        return np.zeros(self._D)

    @property
    def params(self):
        return dict(self._params)

    @property
    def num_params(self) -> int:
        return int(sum(v.numel() for v in self._params.values()))

    def __hash__(self):
        return id(self)

    # -- C-ABI descriptor -----------------------------------------------------------------------
    def host_arrays(self):
        """(family, D, prior_kind, K, prior_a, prior_b, params) as host numpy: what NsModelDesc packs."""
        return (self.log_likelihood.family, self._D, self._prior_kind, self.log_likelihood.K, self._a, self._b,
                self._params_host)

    def desc(self, external: bool = False) -> _lib.NsModelDesc:
        """`external`: describe the model as caller-evaluated (family EXTERNAL) whatever its likelihood is -- how a
        registered family runs through the split slice step when the sampler needs gradients."""
        _lib.require_cuda()
        if self._dev is None:
            a = torch.from_numpy(self._a).cuda()
            b = torch.from_numpy(self._b).cuda()
            p = torch.from_numpy(self._params_host if self._params_host.size else np.zeros(1)).cuda()
            self._dev = (a, b, p)
        a, b, p = self._dev
        if external and not self.is_external:
            from jaxns_b200 import _consts
            return _lib.NsModelDesc(_consts.FAM_EXTERNAL, self._D, self._prior_kind, 0, a.data_ptr(), b.data_ptr(),
                                    p.data_ptr(), 0)
        return _lib.NsModelDesc(self.log_likelihood.family, self._D, self._prior_kind, self.log_likelihood.K,
                                a.data_ptr(), b.data_ptr(), p.data_ptr(), int(self._params_host.size))

    # -- reference methods ----------------------------------------------------------------------
    def call_likelihood(self, X: torch.Tensor) -> torch.Tensor:
        """External likelihood at transformed points X [n, D] -> log L [n] (float64, contiguous, NaN -> -inf as
        framework/ops.py:323-325); the callable sees one [n, size] tensor per variable prior_model returns."""
        out = self.log_likelihood.fn(*[X[:, a:b] for a, b in self._ret_slices])
        out = torch.as_tensor(out, dtype=torch.float64, device=X.device).reshape(-1)
        if out.numel() != X.shape[0]:
            raise ValueError(f"log_likelihood must return one value per row: got {out.numel()} for {X.shape[0]} rows")
        return torch.nan_to_num(out, nan=-float("inf"), posinf=float("inf"), neginf=-float("inf")).contiguous()

    def _forward_batch(self, U: torch.Tensor, want_X: bool, want_L: bool = True):
        _lib.require_cuda()
        U = torch.as_tensor(U, dtype=torch.float64, device="cuda")
        batched = U.dim() == 2
        U2 = U.reshape(-1, self._D).contiguous()
        n = U2.shape[0]
        if self.is_general:
            logL = self._general_log_likelihood(U2) if want_L else None
            if not batched:
                return (logL[0] if want_L else None), None
            return logL, None
        d = self.desc()
        if self.is_external:
            X = torch.empty_like(U2)
            _lib.check(_lib.lib().nsb200_transform_batch(ctypes.byref(d), _lib.ptr(U2), ctypes.c_int64(n), _lib.ptr(X),
                                                          _lib.stream_arg()))
            logL = self.call_likelihood(X) if want_L else None
            if not batched:
                return (logL[0] if want_L else None), (X[0] if want_X else None)
            return logL, (X if want_X else None)
        logL = torch.empty(n, dtype=torch.float64, device="cuda")
        X = torch.empty_like(U2) if want_X else None
        _lib.check(_lib.lib().nsb200_forward_batch(ctypes.byref(d), _lib.ptr(U2), ctypes.c_int64(n), _lib.ptr(logL),
                                                    _lib.ptr(X), _lib.stream_arg()))
        if not batched:
            return logL[0], (X[0] if want_X else None)
        return logL, X

    def forward(self, U, allow_nan: bool = False):
        """log L at U (framework/model.py:167-176); accepts [D] or a batch [n, D] (= vmap(forward))."""
        if allow_nan:
            raise NotImplementedError("allow_nan=True is not supported by the fused kernel (NaN -> -inf).")
        return self._forward_batch(U, False)[0]

    def log_prob_likelihood(self, U, allow_nan: bool = False):
        return self.forward(U, allow_nan)

    def transform(self, U):
        """U -> X dict keyed by prior names (framework/model.py:155-159)."""
        if self.is_general:
            U = torch.as_tensor(U, dtype=torch.float64, device="cuda")
            _, named, _, _ = self._run_generator(U.reshape(-1, self._D))
            return named if U.dim() == 2 else {k: v[0] for k, v in named.items()}
        X = self._forward_batch(U, True, False)[1]
        out, o = {}, 0
        for i, p in enumerate(self._priors):
            n = p.dist.event_size()
            if p.name is not None:
                out[p.name] = X[..., o:o + n]
            o += n
        return out

    def transform_parametrised(self, U):
        """framework/model.py:161-165: the values of the parametrised (singular) priors, keyed by prior name."""
        if not self.is_general:
            return {}
        U = torch.as_tensor(U, dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        out = {}
        with torch.no_grad():
            self._run_generator(U.reshape(-1, self._D), parametrised=out)
        return out if U.dim() == 2 else {k: v[0] for k, v in out.items()}

    def prepare_input(self, U):
        if self.is_general:
            U = torch.as_tensor(U, dtype=torch.float64, device="cuda")
            ret = self._run_generator(U.reshape(-1, self._D))[0]
            return ret if U.dim() == 2 else tuple(r[0] for r in ret)
        return (self._forward_batch(U, True, False)[1],)

    def sample_U(self, key):
        """uniform(split(key, 2)[1], (D,)) (framework/model.py:122-138, context.py:107-109)."""
        from jaxns_b200 import random
        return random.uniform(random.split(key, 2)[1], self._D)

    def log_prob_prior(self, U):
        """Prior log density of the transformed point (framework/model.py:178-187), evaluated on the device
        (post-processing for NestedSamplerResults.log_posterior_density; small torch reductions)."""
        import math
        if self.is_general:
            U = torch.as_tensor(U, dtype=torch.float64, device="cuda")
            lp = self._run_generator(U.reshape(-1, self._D))[2]
            return lp if U.dim() == 2 else lp[0]
        X = self._forward_batch(U, True, False)[1]
        a = self._dev[0]
        b = self._dev[1]
        if self._prior_kind == distributions.Uniform.prior_kind:
            inside = (X >= a) & (X <= a + b)
            lp = torch.where(inside, -torch.log(b).expand_as(X), torch.full_like(X, -math.inf))
        else:
            z = (X - a) / b
            lp = -0.5 * z * z - torch.log(b) - 0.5 * math.log(2.0 * math.pi)
        return lp.sum(dim=-1)

    def sanity_check(self, key, S: int):
        from jaxns_b200 import random
        U = random.uniform(key, S * self._D).reshape(S, self._D)
        logL = self.forward(U)
        if torch.isnan(logL).any():
            raise AssertionError("NaN log-likelihood in sanity check")
        return True

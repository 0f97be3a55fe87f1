"""
Minimal stand-in for the tensorflow_probability distributions the reference's Prior wraps
(/root/reference/src/jaxns/framework/wrapped_tfp_distribution.py:54-84): only what the prior
transform needs (event size + quantile parameters).  Real tfp objects are accepted too when they
expose the same attributes (low/high, loc/scale, loc/scale_tril).
"""
import math

import numpy as np

from jaxns_b200 import _consts


def _is_tensor(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _needs_general(*params) -> bool:
    """A parameter that is a device tensor, or a placeholder of another prior's value (dependent priors), cannot be
    packed into the static per-dimension quantile arrays: the model then evaluates its prior transform as batched
    torch code (framework.Model general mode)."""
    return any(_is_tensor(p) or getattr(p, "_nsb200_placeholder", False) for p in params)


_const_cache = {}


def _t(x, like=None):
    """parameter -> float64 tensor on the device of `like`.  Host constants (numbers, numpy arrays) are uploaded once
    and cached by content: a prior model rebuilds its distributions on every pass, and a fresh host-to-device copy per
    pass would both cost a synchronisation and make the pass impossible to capture in a CUDA graph."""
    import torch
    dev = like.device if like is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if _is_tensor(x):
        return x.to(dtype=torch.float64, device=dev)
    arr = np.asarray(x, np.float64)
    if arr.size > 4096:
        return torch.as_tensor(arr, dtype=torch.float64, device=dev)
    key = (arr.tobytes(), arr.shape, str(dev))
    t = _const_cache.get(key)
    if t is None:
        if len(_const_cache) >= 1024:
            _const_cache.clear()
        t = _const_cache[key] = torch.as_tensor(arr, dtype=torch.float64, device=dev)
    return t


def _rows(x, like):
    """parameter -> [1 or n, size] so that it broadcasts against a batch of draws [n, size]"""
    x = _t(x, like)
    if x.dim() == 0:
        return x.reshape(1, 1)
    if x.dim() == 1:
        return x.reshape(1, -1)
    return x


def _event_sum(lp, X):
    """per-dimension log densities -> one value per row, whichever of the parameters / the points carries the batch"""
    import torch
    return torch.broadcast_tensors(lp, X)[0].sum(dim=-1)


class Distribution:
    prior_kind: int
    dynamic = False  # parameters are tensors / other priors' values: only the torch path below can evaluate it

    def event_size(self) -> int:
        return int(self._size)

    def quantile_params(self):
        """(prior_kind, a[D], b[D]) with X = a + b * U (uniform) or a + b * ndtri(U) (normal)."""
        if self.dynamic or getattr(self, "_a", None) is None:
            raise NotImplementedError("this prior has no static per-dimension quantile parameters")
        return self.prior_kind, self._a, self._b

    # batched torch evaluation (general models, post-processing): U [n, size] -> X [n, size]; X -> log density [n]
    def quantile_torch(self, U):
        raise NotImplementedError

    def log_prob_torch(self, X):
        raise NotImplementedError

    def log_prob(self, x):
        """tfpd's log_prob for use inside torch likelihoods (e.g. Normal(y, sigma).log_prob(0.) with batched y, sigma
        [n, 1]): one value per row."""
        x = _t(x)
        return self.log_prob_torch(x.reshape(1, -1) if x.dim() < 2 else x)


class Uniform(Distribution):
    """tfpd.Uniform(low, high): quantile(u) = u * (high - low) + low."""
    prior_kind = _consts.PRIOR_UNIFORM

    def __init__(self, low=0.0, high=1.0):
        if _needs_general(low, high):
            self.dynamic, self.low, self.high, self._a = True, low, high, None
            self._size = max(1, *[int(p.shape[-1]) if _is_tensor(p) and p.dim() > 0 else int(np.size(p)) if not _is_tensor(p) else 1
                                  for p in (low, high)])
            return
        low, high = np.broadcast_arrays(np.asarray(low, np.float64), np.asarray(high, np.float64))
        self.low = np.atleast_1d(low).reshape(-1).copy()
        self.high = np.atleast_1d(high).reshape(-1).copy()
        self._a = self.low
        self._b = self.high - self.low
        self._size = self._a.size

    def quantile_torch(self, U):
        low, high = _rows(self.low, U), _rows(self.high, U)
        return low + (high - low) * U

    def log_prob_torch(self, X):
        import torch
        low, high = _rows(self.low, X), _rows(self.high, X)
        inside = (X >= low) & (X <= high)
        lp = torch.where(inside, -torch.log(high - low), torch.full_like(X, -math.inf))
        return _event_sum(lp, X)

    def log_prob(self, x):
        if self.dynamic or _is_tensor(x):
            return Distribution.log_prob(self, x)
        x = np.asarray(x, np.float64)
        inside = (x >= self.low) & (x <= self.high)
        return np.sum(np.where(inside, -np.log(self._b), -np.inf), axis=-1)


class Normal(Distribution):
    """tfpd.Normal(loc, scale): quantile(u) = ndtri(u) * scale + loc."""
    prior_kind = _consts.PRIOR_NORMAL

    def __init__(self, loc=0.0, scale=1.0):
        if _needs_general(loc, scale):
            self.dynamic, self.loc, self.scale, self._a = True, loc, scale, None
            self._size = max(1, *[int(p.shape[-1]) if _is_tensor(p) and p.dim() > 0 else int(np.size(p)) if not _is_tensor(p) else 1
                                  for p in (loc, scale)])
            return
        loc, scale = np.broadcast_arrays(np.asarray(loc, np.float64), np.asarray(scale, np.float64))
        self.loc = np.atleast_1d(loc).reshape(-1).copy()
        self.scale = np.atleast_1d(scale).reshape(-1).copy()
        self._a = self.loc
        self._b = self.scale
        self._size = self._a.size

    def quantile_torch(self, U):
        import torch
        return _rows(self.loc, U) + _rows(self.scale, U) * torch.special.ndtri(U)

    def log_prob_torch(self, X):
        import torch
        loc, scale = _rows(self.loc, X), _rows(self.scale, X)
        z = (X - loc) / scale
        return _event_sum((-0.5 * z * z - torch.log(scale) - 0.5 * math.log(2.0 * math.pi)), X)

    def log_prob(self, x):
        if self.dynamic or _is_tensor(x):
            return Distribution.log_prob(self, x)
        z = (np.asarray(x, np.float64) - self.loc) / self.scale
        return np.sum(-0.5 * z * z - np.log(self.scale) - 0.5 * np.log(2 * np.pi), axis=-1)


class MultivariateNormalDiag(Normal):
    def __init__(self, loc, scale_diag):
        super().__init__(loc, scale_diag)


class MultivariateNormalTriL(Normal):
    """tfpd.MultivariateNormalTriL(loc, scale_tril).  The reference decomposes it into
    Sample(Normal) + a TriL bijector (framework/tests/test_prior.py:93-103); only diagonal
    scale_tril (every config of BASELINE.json) maps onto the per-dimension quantile kernel."""

    def __init__(self, loc, scale_tril):
        scale_tril = np.asarray(scale_tril, np.float64)
        if np.any(np.abs(scale_tril - np.diag(np.diag(scale_tril))) > 0):
            # a dense factor is a matrix-vector product inside the prior transform: X = loc + L ndtri(U) (the reference's
            # Sample(Normal) + TriL bijector, framework/tests/test_prior.py:93-103) -- evaluated by the torch path
            self.dynamic, self._a = True, None
            self.loc = np.atleast_1d(np.asarray(loc, np.float64)).reshape(-1)
            self.scale_tril = np.tril(scale_tril)
            self._size = self.scale_tril.shape[0]
            return
        super().__init__(loc, np.diag(scale_tril))
        self.scale_tril = scale_tril

    def quantile_torch(self, U):
        import torch
        if not self.dynamic:
            return super().quantile_torch(U)
        return _rows(self.loc, U) + torch.special.ndtri(U) @ _t(self.scale_tril, U).T

    def log_prob_torch(self, X):
        import torch
        if not self.dynamic:
            return super().log_prob_torch(X)
        L = _t(self.scale_tril, X)
        z = torch.linalg.solve_triangular(L, (X - _rows(self.loc, X)).T, upper=False).T
        return -0.5 * (z * z).sum(dim=-1) - torch.log(torch.diagonal(L)).sum() - 0.5 * L.shape[0] * math.log(2.0 * math.pi)


class MultivariateNormalFullCovariance(MultivariateNormalTriL):
    def __init__(self, loc, covariance_matrix):
        super().__init__(loc, np.linalg.cholesky(np.asarray(covariance_matrix, np.float64)))


class Constant(Distribution):
    """Prior(value): a SingularPrior of the reference (framework/prior.py) -- no U dimensions, the value itself."""
    dynamic = True
    prior_kind = None

    def __init__(self, value):
        self.value = value
        self._size = 0
        self._a = None

    def quantile_torch(self, U):
        v = _t(self.value, U)
        v = v.reshape(1, -1) if v.dim() < 2 else v
        return v.expand(U.shape[0], v.shape[-1]) if v.shape[0] == 1 else v

    def log_prob_torch(self, X):
        import torch
        return torch.zeros(X.shape[0], dtype=torch.float64, device=X.device)


class _QuantileDistribution(Distribution):
    """Scalar families with a closed-form quantile, evaluated by the batched torch prior transform only (models with a
    callable likelihood).  Parameters broadcast over the event: numbers, arrays, tensors or other priors' values."""
    dynamic = True
    prior_kind = None
    _a = None

    def _set(self, **params):
        self._p = params
        self._size = max([1] + [int(p.shape[-1]) if _is_tensor(p) and p.dim() > 0 else (1 if _is_tensor(p) or getattr(p, "_nsb200_placeholder", False) else int(np.size(p)))
                                for p in params.values()])

    def _get(self, name, like):
        return _rows(self._p[name], like)


class Exponential(_QuantileDistribution):
    """tfpd.Exponential(rate): quantile(u) = -log1p(-u) / rate."""

    def __init__(self, rate=1.0):
        self._set(rate=rate)

    def quantile_torch(self, U):
        import torch
        return -torch.log1p(-U) / self._get("rate", U)

    def log_prob_torch(self, X):
        import torch
        rate = self._get("rate", X)
        lp = torch.where(X >= 0, torch.log(rate) - rate * X, torch.full_like(X, -math.inf))
        return _event_sum(lp, X)


class HalfNormal(_QuantileDistribution):
    """tfpd.HalfNormal(scale): quantile(u) = scale * ndtri((1 + u) / 2)."""

    def __init__(self, scale=1.0):
        self._set(scale=scale)

    def quantile_torch(self, U):
        import torch
        return self._get("scale", U) * torch.special.ndtri(0.5 * (1.0 + U))

    def log_prob_torch(self, X):
        import torch
        scale = self._get("scale", X)
        z = X / scale
        lp = 0.5 * math.log(2.0 / math.pi) - torch.log(scale) - 0.5 * z * z
        return _event_sum(torch.where(X >= 0, lp, torch.full_like(lp, -math.inf)), X)


class Cauchy(_QuantileDistribution):
    """tfpd.Cauchy(loc, scale): quantile(u) = loc + scale * tan(pi (u - 1/2))."""

    def __init__(self, loc=0.0, scale=1.0):
        self._set(loc=loc, scale=scale)

    def quantile_torch(self, U):
        import torch
        return self._get("loc", U) + self._get("scale", U) * torch.tan(math.pi * (U - 0.5))

    def log_prob_torch(self, X):
        import torch
        scale = self._get("scale", X)
        z = (X - self._get("loc", X)) / scale
        return _event_sum((-math.log(math.pi) - torch.log(scale) - torch.log1p(z * z)), X)


class HalfCauchy(_QuantileDistribution):
    """tfpd.HalfCauchy(loc, scale): quantile(u) = loc + scale * tan(pi u / 2)."""

    def __init__(self, loc=0.0, scale=1.0):
        self._set(loc=loc, scale=scale)

    def quantile_torch(self, U):
        import torch
        return self._get("loc", U) + self._get("scale", U) * torch.tan(0.5 * math.pi * U)

    def log_prob_torch(self, X):
        import torch
        scale, loc = self._get("scale", X), self._get("loc", X)
        z = (X - loc) / scale
        lp = math.log(2.0 / math.pi) - torch.log(scale) - torch.log1p(z * z)
        return _event_sum(torch.where(X >= loc, lp, torch.full_like(lp, -math.inf)), X)


class Laplace(_QuantileDistribution):
    """tfpd.Laplace(loc, scale): quantile(u) = loc - scale sign(u - 1/2) log(1 - 2 |u - 1/2|)."""

    def __init__(self, loc=0.0, scale=1.0):
        self._set(loc=loc, scale=scale)

    def quantile_torch(self, U):
        import torch
        q = U - 0.5
        return self._get("loc", U) - self._get("scale", U) * torch.sign(q) * torch.log1p(-2.0 * torch.abs(q))

    def log_prob_torch(self, X):
        import torch
        scale = self._get("scale", X)
        return _event_sum((-math.log(2.0) - torch.log(scale) - torch.abs(X - self._get("loc", X)) / scale), X)


class Gumbel(_QuantileDistribution):
    """tfpd.Gumbel(loc, scale): quantile(u) = loc - scale log(-log u)."""

    def __init__(self, loc=0.0, scale=1.0):
        self._set(loc=loc, scale=scale)

    def quantile_torch(self, U):
        import torch
        return self._get("loc", U) - self._get("scale", U) * torch.log(-torch.log(U))

    def log_prob_torch(self, X):
        import torch
        scale = self._get("scale", X)
        z = (X - self._get("loc", X)) / scale
        return _event_sum((-(z + torch.exp(-z)) - torch.log(scale)), X)


class Kumaraswamy(_QuantileDistribution):
    """tfpd.Kumaraswamy(concentration1=a, concentration0=b): quantile(u) = (1 - (1 - u)^(1/b))^(1/a)."""

    def __init__(self, concentration1=1.0, concentration0=1.0):
        self._set(a=concentration1, b=concentration0)

    def quantile_torch(self, U):
        import torch
        a, b = self._get("a", U), self._get("b", U)
        return torch.exp(torch.log(-torch.expm1(torch.log1p(-U) / b)) / a)

    def log_prob_torch(self, X):
        import torch
        a, b = self._get("a", X), self._get("b", X)
        lp = torch.log(a) + torch.log(b) + (a - 1.0) * torch.log(X) + (b - 1.0) * torch.log1p(-X ** a)
        inside = (X >= 0) & (X <= 1)
        return _event_sum(torch.where(inside, lp, torch.full_like(lp, -math.inf)), X)


class TruncatedNormal(_QuantileDistribution):
    """tfpd.TruncatedNormal(loc, scale, low, high): quantile(u) = loc + scale ndtri(Phi(a) + u (Phi(b) - Phi(a)))."""

    def __init__(self, loc=0.0, scale=1.0, low=-1.0, high=1.0):
        self._set(loc=loc, scale=scale, low=low, high=high)

    def _cdf_bounds(self, like):
        import torch
        loc, scale = self._get("loc", like), self._get("scale", like)
        ca = torch.special.ndtr((self._get("low", like) - loc) / scale)
        cb = torch.special.ndtr((self._get("high", like) - loc) / scale)
        return loc, scale, ca, cb

    def quantile_torch(self, U):
        import torch
        loc, scale, ca, cb = self._cdf_bounds(U)
        x = loc + scale * torch.special.ndtri(ca + U * (cb - ca))
        return torch.minimum(torch.maximum(x, self._get("low", U)), self._get("high", U))

    def log_prob_torch(self, X):
        import torch
        loc, scale, ca, cb = self._cdf_bounds(X)
        z = (X - loc) / scale
        lp = -0.5 * z * z - 0.5 * math.log(2.0 * math.pi) - torch.log(scale) - torch.log(cb - ca)
        inside = (X >= self._get("low", X)) & (X <= self._get("high", X))
        return _event_sum(torch.where(inside, lp, torch.full_like(lp, -math.inf)), X)


def from_any(dist) -> Distribution:
    """Accept our shim or a duck-typed tfp distribution."""
    if isinstance(dist, Distribution):
        return dist
    name = type(dist).__name__
    if hasattr(dist, "low") and hasattr(dist, "high"):
        return Uniform(np.asarray(dist.low), np.asarray(dist.high))
    if hasattr(dist, "scale_tril"):
        return MultivariateNormalTriL(np.asarray(dist.loc), np.asarray(dist.scale_tril))
    if hasattr(dist, "loc") and hasattr(dist, "scale"):
        return Normal(np.asarray(dist.loc), np.asarray(dist.scale))
    if _is_tensor(dist) or isinstance(dist, (int, float, np.ndarray, list, tuple)):
        return Constant(dist)
    raise NotImplementedError(f"Unsupported prior distribution {name}")

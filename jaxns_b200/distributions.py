"""
Minimal stand-in for the tensorflow_probability distributions the reference's Prior wraps
(/root/reference/src/jaxns/framework/wrapped_tfp_distribution.py:54-84): only what the prior
transform needs (event size + quantile parameters).  Real tfp objects are accepted too when they
expose the same attributes (low/high, loc/scale, loc/scale_tril).
"""
import numpy as np

from jaxns_b200 import _consts


class Distribution:
    prior_kind: int

    def event_size(self) -> int:
        return int(self._a.size)

    def quantile_params(self):
        """(prior_kind, a[D], b[D]) with X = a + b * U (uniform) or a + b * ndtri(U) (normal)."""
        return self.prior_kind, self._a, self._b


class Uniform(Distribution):
    """tfpd.Uniform(low, high): quantile(u) = u * (high - low) + low."""
    prior_kind = _consts.PRIOR_UNIFORM

    def __init__(self, low=0.0, high=1.0):
        low, high = np.broadcast_arrays(np.asarray(low, np.float64), np.asarray(high, np.float64))
        self.low = np.atleast_1d(low).reshape(-1).copy()
        self.high = np.atleast_1d(high).reshape(-1).copy()
        self._a = self.low
        self._b = self.high - self.low

    def log_prob(self, x):
        x = np.asarray(x, np.float64)
        inside = (x >= self.low) & (x <= self.high)
        return np.sum(np.where(inside, -np.log(self._b), -np.inf), axis=-1)


class Normal(Distribution):
    """tfpd.Normal(loc, scale): quantile(u) = ndtri(u) * scale + loc."""
    prior_kind = _consts.PRIOR_NORMAL

    def __init__(self, loc=0.0, scale=1.0):
        loc, scale = np.broadcast_arrays(np.asarray(loc, np.float64), np.asarray(scale, np.float64))
        self.loc = np.atleast_1d(loc).reshape(-1).copy()
        self.scale = np.atleast_1d(scale).reshape(-1).copy()
        self._a = self.loc
        self._b = self.scale

    def log_prob(self, x):
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
            raise NotImplementedError("MultivariateNormalTriL prior with a non-diagonal scale_tril is not "
                                      "supported by the fused prior transform yet.")
        super().__init__(loc, np.diag(scale_tril))


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
    raise NotImplementedError(f"Unsupported prior distribution {name}")

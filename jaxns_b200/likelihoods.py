"""
Registered likelihood families: objects a user passes as `log_likelihood` to Model so that the fused
slice kernel can evaluate them as device functions (include/nsb200.h NSB200_FAM_*).  Each mirrors a
likelihood used by the reference's benchmarks / examples:

* DenseGaussianLikelihood   tfpd.MultivariateNormalTriL(loc, chol(cov)).log_prob
                            (/root/reference/docs/papers/phantom-powered-nested-sampling/run_experiment.py:23-50,
                             src/jaxns/tests/conftest.py:217-251)
* GaussianMixtureLikelihood logaddexp of diagonal Gaussians (benchmarks/difficult_problems/main.py:95-125)
* EggBoxLikelihood          (2 + prod cos(theta/2))^5 (docs/examples/egg_box.ipynb cell 2)
* RosenbrockLikelihood      benchmarks/difficult_problems/main.py:69-92
* GaussianShellsLikelihood  docs/examples/gaussian_shells.ipynb cell 2
"""
import numpy as np

from jaxns_b200 import _consts


class RegisteredLikelihood:
    """Base class; `family` and `pack(D)` define the C-ABI model descriptor."""
    family: int
    K: int = 0
    __nsb200_family__ = True

    def _dev(self, name: str, array: np.ndarray, like):
        """A host parameter as a device tensor, uploaded once per device (the gradient rounds are captured in a CUDA
        graph, which rules out host-to-device copies inside them)."""
        import torch
        cache = self.__dict__.setdefault("_dev_cache", {})
        key = (name, str(like.device))
        if key not in cache:
            if callable(array):
                array = array()
            cache[key] = torch.as_tensor(np.ascontiguousarray(array, np.float64), device=like.device)
        return cache[key]

    def pack(self, D: int) -> np.ndarray:
        return np.zeros(0)

    def __call__(self, *args):
        raise RuntimeError("Registered likelihoods are evaluated on the device; use Model.forward(U).")

    def log_prob_torch(self, X):
        """The family as differentiable torch code over X [n, D]: only the gradient variants of the slice sampler
        use it (Model.grad_U, the analogue of jax.grad(model.forward), uni_slice_sampler.py:135); values always come
        from the kernels."""
        raise NotImplementedError


class ExternalLikelihood(RegisteredLikelihood):
    """An arbitrary user likelihood: a BATCHED device callable `fn(*variables) -> log_L[n]` over torch CUDA
    float64 tensors (each variable [n, size], in the order prior_model returns them) -- the analogue of the
    reference's vmap(log_likelihood) compiled by XLA (framework/ops.py:302-326).  The slice step is split
    into propose / accept kernels around this call (include/nsb200.h nsb200_split_*); nothing runs on the CPU.
    Model wraps any plain callable in this class."""
    family = _consts.FAM_EXTERNAL

    def __init__(self, fn):
        if not callable(fn):
            raise TypeError("log_likelihood must be a RegisteredLikelihood or a callable")
        self.fn = fn

    def __call__(self, *args):
        return self.fn(*args)


class DenseGaussianLikelihood(RegisteredLikelihood):
    family = _consts.FAM_GAUSS_DENSE

    def __init__(self, loc, covariance_matrix=None, scale_tril=None):
        self.loc = np.atleast_1d(np.asarray(loc, np.float64))
        if (covariance_matrix is None) == (scale_tril is None):
            raise ValueError("Give exactly one of covariance_matrix / scale_tril.")
        if scale_tril is None:
            scale_tril = np.linalg.cholesky(np.asarray(covariance_matrix, np.float64))
        self.scale_tril = np.tril(np.asarray(scale_tril, np.float64))

    def pack(self, D: int) -> np.ndarray:
        if self.loc.size != D or self.scale_tril.shape != (D, D):
            raise ValueError(f"Gaussian likelihood is {self.loc.size}-D but the prior has {D} dims.")
        L = self.scale_tril
        Linv = np.tril(np.linalg.solve(L, np.eye(D)))
        c = -np.sum(np.log(np.diag(L))) - 0.5 * D * np.log(2.0 * np.pi)
        return np.concatenate([[c], self.loc, Linv.reshape(-1)])

    def log_prob_torch(self, X):
        import torch
        D = self.loc.size
        Linv = self._dev("Linv", lambda: np.tril(np.linalg.solve(self.scale_tril, np.eye(D))), X)
        z = (X - self._dev("loc", self.loc, X)) @ Linv.T
        c = -np.sum(np.log(np.diag(self.scale_tril))) - 0.5 * D * np.log(2.0 * np.pi)
        return c - 0.5 * (z * z).sum(-1)


class GaussianMixtureLikelihood(RegisteredLikelihood):
    """log sum_k w_k N(x | mean_k, diag(var_k)); weights default to 1 (the reference's spike-and-slab
    adds two normalised densities)."""
    family = _consts.FAM_GAUSS_MIX_DIAG

    def __init__(self, means, variances, log_weights=None):
        self.means = np.atleast_2d(np.asarray(means, np.float64))
        self.K, D = self.means.shape
        v = np.asarray(variances, np.float64)
        if v.ndim == 1 and v.size == self.K:
            v = np.repeat(v[:, None], D, axis=1)
        self.variances = np.broadcast_to(v, (self.K, D)).copy()
        self.log_weights = np.zeros(self.K) if log_weights is None else np.asarray(log_weights, np.float64)

    def pack(self, D: int) -> np.ndarray:
        if self.means.shape[1] != D:
            raise ValueError(f"Mixture is {self.means.shape[1]}-D but the prior has {D} dims.")
        rows = []
        for k in range(self.K):
            logc = self.log_weights[k] - 0.5 * np.sum(np.log(2.0 * np.pi * self.variances[k]))
            rows.append(np.concatenate([[logc], self.means[k], 1.0 / np.sqrt(self.variances[k])]))
        return np.concatenate(rows)

    def log_prob_torch(self, X):
        import torch
        mu = self._dev("means", self.means, X)
        var = self._dev("variances", self.variances, X)
        logc = self._dev("logc", lambda: self.log_weights - 0.5 * np.sum(np.log(2.0 * np.pi * self.variances), axis=1), X)
        q = ((X[:, None, :] - mu[None]) ** 2 / var[None]).sum(-1)
        return torch.logsumexp(logc[None] - 0.5 * q, dim=1)


class EggBoxLikelihood(RegisteredLikelihood):
    family = _consts.FAM_EGGBOX

    def log_prob_torch(self, X):
        import torch
        return (2.0 + torch.cos(0.5 * X).prod(-1)) ** 5


class RosenbrockLikelihood(RegisteredLikelihood):
    family = _consts.FAM_ROSENBROCK

    def log_prob_torch(self, X):
        a = X[:, 1:] - X[:, :-1] ** 2
        b = 1.0 - X[:, :-1]
        return -(100.0 * a * a + b * b).sum(-1)


class GaussianShellsLikelihood(RegisteredLikelihood):
    family = _consts.FAM_SHELLS

    def __init__(self, centres, radii, widths):
        self.centres = np.atleast_2d(np.asarray(centres, np.float64))
        self.K = self.centres.shape[0]
        self.radii = np.broadcast_to(np.asarray(radii, np.float64), (self.K,)).copy()
        self.widths = np.broadcast_to(np.asarray(widths, np.float64), (self.K,)).copy()

    def pack(self, D: int) -> np.ndarray:
        if self.centres.shape[1] != D:
            raise ValueError(f"Shells are {self.centres.shape[1]}-D but the prior has {D} dims.")
        return np.concatenate([np.concatenate([[self.widths[k], self.radii[k]], self.centres[k]])
                               for k in range(self.K)])

    def log_prob_torch(self, X):
        import torch
        c = self._dev("centres", self.centres, X)
        w = self._dev("widths", self.widths, X)
        r = self._dev("radii", self.radii, X)
        e = torch.sqrt(((X[:, None, :] - c[None]) ** 2).sum(-1)) - r[None]
        g = -0.5 * e * e / (w * w)[None] - torch.log(torch.sqrt(2.0 * np.pi * w * w))[None]
        return torch.logsumexp(g, dim=1)


def jaxify_likelihood(log_likelihood, vectorised: bool = False):
    """framework/jaxify.py:15-53: wraps a host (numpy) log-likelihood so that Model accepts it.  The reference routes
    it through jax.pure_callback; here the split slice step hands the transformed batch to the host function and
    takes the values back to the device.  `vectorised`: the function handles a leading batch dimension itself."""
    import warnings
    import torch
    warnings.warn(
        "You're using a non-JAX log-likelihood function. This may be slower than a JAX log-likelihood function. "
        "Also, you are responsible for ensuring that the function is deterministic. "
        "Also, you cannot use learnable parameters in the likelihood call."
    )

    def _log_likelihood(*args):
        host = [a.detach().cpu().numpy() for a in args]
        if vectorised:
            out = np.asarray(log_likelihood(*host), np.float64)
        else:
            out = np.asarray([log_likelihood(*[h[i] for h in host]) for i in range(host[0].shape[0])], np.float64)
        return torch.from_numpy(out.reshape(-1)).to(args[0].device)

    _log_likelihood._nsb200_host_callback = True  # leaves the device: the likelihood rounds are not graph-captured
    return _log_likelihood

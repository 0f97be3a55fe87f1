"""North-star correctness check: |logZ - analytic| < 3 sigma over 10 seeds (PRNGKey(0..9)) for the Gaussian
configs, and the mean error against sigma/sqrt(10).  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jaxns_b200 as j
from jaxns_b200 import distributions as tfpd, likelihoods as lk, random


def gauss(D):
    cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)
    def prior_model():
        x = yield j.Prior(tfpd.MultivariateNormalTriL(loc=np.zeros(D), scale_tril=np.eye(D)), name="x")
        return x
    S = cov + np.eye(D)
    L = np.linalg.cholesky(S)
    z = np.linalg.solve(L, np.full(D, 15.0))
    true = float(-0.5 * z @ z - np.log(np.diag(L)).sum() - 0.5 * D * np.log(2 * np.pi))
    return j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov)), true


for D, N in ((2, 500), (8, 240), (32, 3200)):
    model, true = gauss(D)
    ns = j.NestedSampler(model=model, num_live_points=N)
    errs, sig, ev = [], [], []
    for seed in range(10):
        reason, state = ns(random.PRNGKey(seed))
        r = ns.to_results(reason, state)
        errs.append(r.log_Z_mean - true); sig.append(r.log_Z_uncert); ev.append(r.total_num_likelihood_evaluations)
    errs, sig = np.array(errs), np.array(sig)
    print(f"D={D:3d} N={N:5d} analytic {true:.4f}: errors {np.round(errs, 3).tolist()} sigma {sig.mean():.3f} "
          f"max|err|/sigma {np.max(np.abs(errs) / sig):.2f} mean err {errs.mean():+.3f} = {errs.mean() / (sig.mean() / np.sqrt(10)):+.2f} sigma_mean "
          f"evals/run {np.mean(ev):.3g}")

#!/usr/bin/env python
"""Throughput of the caller-evaluated (split propose / accept) path: config 2's problem with the likelihood written as a
torch callable instead of the registered family.  Device-timed over the first SHELLS shells."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import distributions as tfpd, likelihoods as lk, random

D, N = 32, 3200
SHELLS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)
Linv = torch.from_numpy(np.linalg.inv(np.linalg.cholesky(cov))).cuda()
mu = torch.full((D,), 15.0, dtype=torch.float64, device="cuda")
c = float(-np.sum(np.log(np.diag(np.linalg.cholesky(cov)))) - 0.5 * D * np.log(2 * np.pi))


def prior_model():
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
    return x


def log_likelihood(x):
    z = (x - mu) @ Linv.T
    return c - 0.5 * (z * z).sum(-1)


for name, model in (("callable", j.Model(prior_model, log_likelihood)),
                    ("registered", j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov)))):
    ns = j.NestedSampler(model=model, num_live_points=N)
    tc = j.TerminationCondition(max_samples=float(SHELLS * N // 2))
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        reason, state = ns(random.PRNGKey(rep), tc)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    evals = int(ns.nested_sampler.last_register.num_likelihood_evaluations)
    ms = e0.elapsed_time(e1)
    print(f"{name:10s}: {SHELLS} shells, {evals} evals in {ms:.1f} ms (wall {1e3 * wall:.1f}) = {evals / ms / 1e3:.2f} M evals/s, "
          f"{ms / SHELLS:.2f} ms per shell")

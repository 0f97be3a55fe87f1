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


from jaxns_b200 import samplers

cases = [("callable", P, g) for P in (1, 2, 4, 8) for g in (0, 1)] + [("registered", 0, 0)]
for name, P, g in cases:
    if name == "callable":
        samplers.SPLIT_PROPOSALS = P
        os.environ["NSB200_SPLIT_GRAPH"] = str(g)
        model = j.Model(prior_model, log_likelihood)
    else:
        model = j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov))
    ns = j.NestedSampler(model=model, num_live_points=N)
    tc = j.TerminationCondition(max_samples=float(SHELLS * N // 2))
    first = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reason, state = ns(random.PRNGKey(rep), tc)
        e1.record()
        torch.cuda.synchronize()
        first = first if first is not None else e0.elapsed_time(e1)
    evals = int(ns.nested_sampler.last_register.num_likelihood_evaluations)
    ms = e0.elapsed_time(e1)
    tag = f"{name} P={P} graph={g}" if name == "callable" else name
    print(f"{tag:24s}: {SHELLS} shells, {evals} evals in {ms:8.1f} ms = {evals / ms / 1e3:7.2f} M evals/s, {ms / SHELLS:6.2f} ms per shell (first run {first:.0f} ms)")

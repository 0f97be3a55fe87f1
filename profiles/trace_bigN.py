#!/usr/bin/env python
"""Device timeline (NSB200_TRACE=1) of a config-2 body at the live-set size of an 8-GPU weak-scaling run
(num_live_points = 25600), all chains on one GPU: what the merge / register-update kernels cost at that size."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import distributions as tfpd, likelihoods as lk, random

D = 32
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25600
cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)


def prior_model():
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
    return x


ns = j.NestedSampler(model=j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov)),
                     num_live_points=N)
for rep in range(2):
    torch.cuda.synchronize()
    reason, state = ns(random.PRNGKey(rep), j.TerminationCondition(max_samples=float(N // 2 * 70)))
    torch.cuda.synchronize()
print("iterations", ns.nested_sampler.last_profile["iterations"])

#!/usr/bin/env python
"""Quick device-timed A/B of config 2 (32-D Gaussian, N=3200): NSB200_LIB=<so> python profiles/quick_cfg2.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, distributions as tfpd, likelihoods as lk, random

D = 32
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)


def prior_model():
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
    return x


ns = j.NestedSampler(model=j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov)),
                     num_live_points=3200)
ms_all, sl_all, ev_all, it_all, lz = [], [], [], [], []
for s in range(-3, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reason, state = ns(random.PRNGKey(max(s, 0)))
    e1.record()
    torch.cuda.synchronize()
    if s < 0:
        continue
    prof = ns.nested_sampler.last_profile
    ms_all.append(e0.elapsed_time(e1))
    sl_all.append(prof["slice_ms"])
    it_all.append(prof["iterations"])
    ev_all.append(int(ns.nested_sampler.last_register.num_likelihood_evaluations))
    if s < 3:
        r = ns.to_results(reason, state)
        lz.append(round(r.log_Z_mean, 3))
print(f"lib={os.path.basename(_lib.so_path())} runs={reps} ms/run={np.mean(ms_all):.2f} (min {np.min(ms_all):.2f}) "
      f"slice_ms/iter={np.sum(sl_all) / np.sum(it_all):.4f} iters={np.mean(it_all):.1f} "
      f"evals/s={np.sum(ev_all) / np.sum(ms_all) * 1e3:.4g} slice_evals/s={np.sum(ev_all) / np.sum(sl_all) * 1e3:.4g} logZ={lz} "
      f"runs_ms={[round(x, 1) for x in ms_all]}")

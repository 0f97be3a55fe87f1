#!/usr/bin/env python
"""Can the chain-stream generator be overtaken?  Config 2 with the loop on a high-priority stream (the generator's
side stream has the default = lowest priority) and the generator cut into short CTAs (NSB200_GEN_TPB / NSB200_GEN_SMS)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, distributions as tfpd, likelihoods as lk, random

D = 32
cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)


def prior_model():
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
    return x


model = j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov))
hi = torch.cuda.Stream(priority=-1)
cases = [("default", False, -1, -1), ("hi-prio loop", True, -1, -1)]
for tpb, ctas in ((128, 1184), (128, 2368), (256, 1184), (256, 592), (512, 296)):
    cases.append((f"hi-prio, gen {ctas} x {tpb}", True, tpb, ctas))
    cases.append((f"lo-prio, gen {ctas} x {tpb}", False, tpb, ctas))
for name, use_hi, tpb, ctas in cases:
    _lib.set_option("NSB200_GEN_TPB", tpb)
    _lib.set_option("NSB200_GEN_SMS", ctas)
    ns = j.NestedSampler(model=model, num_live_points=3200)
    ms = []
    stream = hi if use_hi else torch.cuda.current_stream()
    with torch.cuda.stream(stream):
        for s in range(-2, 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            reason, state = ns(random.PRNGKey(max(s, 0)))
            e1.record()
            torch.cuda.synchronize()
            if s >= 0:
                ms.append(e0.elapsed_time(e1))
        res = ns.to_results(reason, state)
    print(f"{name:28s}: median {np.median(ms):7.2f} ms  {np.round(ms, 1)}  logZ {float(res.log_Z_mean):.3f}", flush=True)
    del ns

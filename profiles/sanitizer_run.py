#!/usr/bin/env python
"""Tiny invocation of every kernel family for compute-sanitizer (racecheck / memcheck):
    compute-sanitizer --tool racecheck python profiles/sanitizer_run.py
slice chains (lane kernel, DMMA kernel, warp team), stream generator, merge (brute + tiled), register update, radix sort,
tree counts, evidence scan, sample_evidence, split propose / accept (plain, gradient_slice, gradient_guided)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, random, utils
from jaxns_b200.internals.tree_structure import SampleTreeGraph, argsort, count_crossed_edges
from tests.models import product_models

for name, D, N in (("gauss", 32, 64), ("gauss", 8, 64), ("eggbox", 2, 4200), ("rosenbrock", 10, 64)):
    for impl in ((0, 1, 2) if name == "gauss" else (0,)):
        _lib.set_option("NSB200_SLICE_MMA", impl)
        model = product_models()[name](D)
        ns = j.NestedSampler(model=model, num_live_points=N, max_samples=N * 8, s=1 if N > 1000 else None,
                             k=2 if name == "rosenbrock" else None)
        reason, state = ns(random.PRNGKey(1), j.TerminationCondition(max_samples=float(N * 4)))
        res = ns.to_results(reason, state)
        print(name, D, N, "impl", impl, "reason", reason, "logZ %.3f" % res.log_Z_mean, flush=True)
_lib.set_option("NSB200_SLICE_MMA", -1)
x = torch.randn(5000, dtype=torch.float64, device="cuda")
assert torch.equal(argsort(x), torch.argsort(x, stable=True))
lz = utils.sample_evidence(random.PRNGKey(2), res.num_live_points_per_sample, res.log_L_samples, S=8)
print("sample_evidence", float(lz.mean()), flush=True)


def ext(x):
    return -0.5 * (x * x).sum(dim=1)


def prior_model():
    from jaxns_b200 import distributions as tfpd
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(3), scale=np.ones(3)), name="x")
    return x


ns = j.NestedSampler(model=j.Model(prior_model, ext), num_live_points=48, max_samples=400)
reason, state = ns(random.PRNGKey(0), j.TerminationCondition(max_samples=200.0))
print("external", reason, flush=True)
# gradient variants (k_split_step mode 2, k_split_export_U0) with the default 4 speculative proposals per round
import warnings
warnings.simplefilter("ignore")
for kw in (dict(gradient_slice=True), dict(gradient_guided=True)):
    sampler = j.UniDimSliceSampler(model=j.Model(prior_model, ext), num_slices=4, num_phantom_save=1, midpoint_shrink=True,
                                   perfect=True, **kw)
    sns = j.ShardedStaticNestedSampler(model=sampler.model, max_samples=400, init_efficiency_threshold=0.1, sampler=sampler,
                                       num_live_points=48)
    reason, reg, state = sns._run(random.PRNGKey(0), j.TerminationCondition(max_samples=200.0))
    print("gradient", kw, reason, flush=True)
torch.cuda.synchronize()
print("sanitizer_run done")

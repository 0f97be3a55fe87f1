#!/usr/bin/env python
"""Lane-per-dimension slice kernel vs the DMMA kernel as the number of chains per GPU grows (32-D correlated
Gaussian, S = 160): one launch of get_samples at a mid-run contour, CUDA-event timed through the B1 entry point."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, random
from jaxns_b200.types import LivePointCollection
from tests.models import product_models

D, S = 32, 160
model = product_models()["gauss"](D)
sampler = j.UniDimSliceSampler(model=model, num_slices=S, num_phantom_save=0, midpoint_shrink=True, perfect=True)
# a mid-run live set: run the engine for 40 shells and take its live points
ns = j.NestedSampler(model=model, num_live_points=3200)
for N in (3200, 6400, 12800, 25600, 51200):
    ns = j.NestedSampler(model=model, num_live_points=N, max_samples=N * 60)
    reason, state = ns(random.PRNGKey(0), j.TerminationCondition(max_samples=float(N * 20)))
    n = min(state.num_samples, ns.nested_sampler.max_samples)
    sc = state.sample_collection
    live_U, live_logL = sc.U_samples[n - N:n].contiguous(), sc.log_L[n - N:n].contiguous()
    order = torch.argsort(live_logL, stable=True)
    st = LivePointCollection(None, live_U[order].contiguous(), None, live_logL[order].contiguous(), None)
    m = N // 2
    contour = float(st.log_L[m - 1].item())
    row = [f"N={N:6d} chains={m:6d}"]
    for impl, P in [(0, 0), (1, 1), (1, 2), (1, 4)]:
        _lib.set_option("NSB200_SLICE_MMA", impl)
        _lib.set_option("NSB200_MMA_P", P)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            sample, _ = sampler.get_samples_batch(random.PRNGKey(5), contour, st, m)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ev = int(sample.num_likelihood_evaluations.sum().item())
        row.append(f"{'lane' if impl == 0 else 'dmma P=%d' % P}: {best:7.3f} ms ({ev / best / 1e6:6.2f} Gevals/s)")
    print(" | ".join(row))
_lib.set_option("NSB200_SLICE_MMA", -1)
_lib.set_option("NSB200_MMA_P", -1)

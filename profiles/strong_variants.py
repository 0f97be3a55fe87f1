#!/usr/bin/env python
"""Strong scaling leaves 1600 / n_gpus chains per GPU: which slice kernel variant is fastest when the GPU is nearly
empty?  One get_samples launch over chains [0, n) of the 1600 of config 2 at a mid-run contour."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, random
from jaxns_b200.types import LivePointCollection
from tests.models import product_models

D, S, N = 32, 160, 3200
model = product_models()["gauss"](D)
sampler = j.UniDimSliceSampler(model=model, num_slices=S, num_phantom_save=0, midpoint_shrink=True, perfect=True)
ns = j.NestedSampler(model=model, num_live_points=N, max_samples=N * 60)
reason, state = ns(random.PRNGKey(0), j.TerminationCondition(max_samples=float(N * 30)))
n = min(state.num_samples, ns.nested_sampler.max_samples)
sc = state.sample_collection
live_U, live_logL = sc.U_samples[n - N:n].contiguous(), sc.log_L[n - N:n].contiguous()
order = torch.argsort(live_logL, stable=True)
st = LivePointCollection(None, live_U[order].contiguous(), None, live_logL[order].contiguous(), None)
m = N // 2
contour = float(st.log_L[m - 1].item())
variants = [("lane P=1", 0, 1, 0), ("lane P=2", 0, 2, 0), ("lane P=4", 0, 4, 0), ("team W=2", 2, 0, 0), ("team W=4", 4, 0, 0),
            ("dmma P=2", 1, 0, 2), ("dmma P=4", 1, 0, 4)]
for nch in (200, 400, 800, 1600):
    row = [f"chains={nch:5d}"]
    for name, impl, spec, P in variants:
        _lib.set_option("NSB200_SLICE_MMA", impl)
        _lib.set_option("NSB200_SPEC", spec if spec else -1)
        _lib.set_option("NSB200_MMA_P", P if P else -1)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            sample, _ = sampler.get_samples_batch(random.PRNGKey(5), contour, st, m, 0, nch)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row.append(f"{name} {best:6.3f}")
    print(" | ".join(row) + "  (ms, incl. stream generation)")
for o in ("NSB200_SLICE_MMA", "NSB200_SPEC", "NSB200_MMA_P"):
    _lib.set_option(o, -1)

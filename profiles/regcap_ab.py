#!/usr/bin/env python
"""A/B of a register-capped build of the lane slice kernel: NSB200_LIB=<so> python profiles/regcap_ab.py
One get_samples launch (B1, incl. stream generation) of the 32-D Gaussian at a mid-run contour for a growing number of
chains, P = 1 / 2 / 4 speculative proposals, then whole config-2 runs (device-timed)."""
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
print("lib:", _lib._SO)
for N in (3200, 12800, 51200):
    ns = j.NestedSampler(model=model, num_live_points=N, max_samples=N * 60)
    reason, state = ns(random.PRNGKey(0), j.TerminationCondition(max_samples=float(N * 20)))
    n = min(state.num_samples, ns.nested_sampler.max_samples)
    sc = state.sample_collection
    live_U, live_logL = sc.U_samples[n - N:n].contiguous(), sc.log_L[n - N:n].contiguous()
    order = torch.argsort(live_logL, stable=True)
    st = LivePointCollection(None, live_U[order].contiguous(), None, live_logL[order].contiguous(), None)
    m = N // 2
    contour = float(st.log_L[m - 1].item())
    for nch in ((200, 1600) if N == 3200 else (m,)):
        row = [f"chains={nch:6d}"]
        for spec in (1, 2, 4):
            _lib.set_option("NSB200_SPEC", spec)
            best = 1e9
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                sample, _ = sampler.get_samples_batch(random.PRNGKey(5), contour, st, m, 0, nch)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            row.append(f"P={spec} {best:7.3f} ms")
        print(" | ".join(row))
    del ns
_lib.set_option("NSB200_SPEC", -1)
ns = j.NestedSampler(model=model, num_live_points=3200)
ms = []
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
print(f"config 2 runs: {np.round(ms, 2)} ms, median {np.median(ms):.2f}; last logZ {float(res.log_Z_mean):.3f}")

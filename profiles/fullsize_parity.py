#!/usr/bin/env python
"""How far do GPU and oracle slice chains agree at the true config sizes?  (Slice sampling is chaotic: a bracket
end is (1 - U_j) / d_j, so rounding-level differences grow per move.)  Prints, per configuration and kernel, the
fraction of chains with identical n_evals and the error of the final points against the oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, random
from jaxns_b200.types import LivePointCollection
from oracle import oracle
from tests.models import product_models, to_oracle

CASES = [("gauss", 32, 3200, 160, 0, True), ("eggbox", 2, 10000, 20, 0, False), ("mixture", 100, 12500, 50, 0, True)]
oracle.set_num_threads(os.cpu_count())
for name, D, N, S, k, midpoint in CASES:
    model = product_models()[name](D)
    om = to_oracle(model, oracle)
    oU, ologL, _ = oracle.init_batch(om, random.PRNGKey(3), N)
    order = np.argsort(ologL, kind="stable")
    live_U, live_logL = oU[order], ologL[order]
    m = N // 2
    contour = live_logL[m - 1]
    key = random.PRNGKey(11)
    t0 = time.perf_counter()
    exp = oracle.slice_batch(om, key, contour, live_U, live_logL, S, k, midpoint, num_samples=m)
    t_or = time.perf_counter() - t0
    sampler = j.UniDimSliceSampler(model=model, num_slices=S, num_phantom_save=k, midpoint_shrink=midpoint, perfect=True)
    state = LivePointCollection(None, torch.from_numpy(live_U).cuda(), None, torch.from_numpy(live_logL).cuda(), None)
    variants = [(0, 0)] + ([(1, 1), (1, 2), (1, 4)] if name == "gauss" else [])
    for impl, P in variants:
        _lib.set_option("NSB200_SLICE_MMA", impl)
        _lib.set_option("NSB200_MMA_P", P)
        sample, _ = sampler.get_samples_batch(key, contour, state, m)
        nev = sample.num_likelihood_evaluations.cpu().numpy()
        U = sample.U_sample.cpu().numpy()
        same = nev == exp["n_evals"]
        err = np.abs(U - exp["U"]).max(axis=1)
        print(f"{name} D={D} N={N} S={S} impl={impl} P={P}: chains={m} n_evals identical {same.mean():.4f} | "
              f"|U - oracle| < 1e-9 on {np.mean(err < 1e-9):.4f}, < 1e-6 on {np.mean(err < 1e-6):.4f} | "
              f"sum n_evals gpu {nev.sum()} oracle {exp['n_evals'].sum()} | oracle {t_or:.1f}s")
    _lib.set_option("NSB200_SLICE_MMA", -1)
    _lib.set_option("NSB200_MMA_P", -1)

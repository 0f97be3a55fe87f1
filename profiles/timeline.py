#!/usr/bin/env python
"""Where the generator and the slice kernel of one body run relative to each other (needs a -DNSB_TIMELINE build:
NSB200_LIB=... python profiles/timeline.py).  Steps the engine body by body and reads %globaltimer stamps."""
import ctypes
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


ns = j.NestedSampler(model=j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov)),
                     num_live_points=3200)
inner = ns.nested_sampler
L = _lib.lib()
eng = inner._make_engine()
st = _lib.stream_arg()
tc = _lib.NsTermCond()
_lib.check(L.nsb200_engine_init(eng.h, _lib.key_arg(random.PRNGKey(0)), ctypes.byref(tc), st))
out = (ctypes.c_ulonglong * 4)()
rows = []
for body in range(70):
    L.nsb200_debug_timeline(out, 1)
    _lib.check(L.nsb200_engine_step(eng.h, st))
    L.nsb200_debug_timeline(out, 0)
    g0, g1, s0, s1 = [int(x) for x in out]
    rows.append((g1 - g0, s0 - g0, s1 - s0, s1 - g0))
r = np.array(rows[20:], dtype=np.float64) / 1e3
print("bodies 20..69 (us): generator duration %.0f | slice start after generator start %.0f | slice duration %.0f | "
      "pair span %.0f" % tuple(r.mean(0)))

#!/usr/bin/env python
"""Max relative error of the device prior quantile against scipy's ndtri (the check of
tests/test_gpu_parity.py::test_prior_quantile_accuracy_vs_scipy, printed instead of asserted), for A/B builds."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.special import ndtri
import jaxns_b200 as j
from jaxns_b200 import _lib, distributions as tfpd, likelihoods as lk
D = 32
def prior_model():
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
    return x
model = j.Model(prior_model, lk.DenseGaussianLikelihood(np.zeros(D), covariance_matrix=np.eye(D)))
rng = np.random.default_rng(0)
n = 1 << 16
U = rng.uniform(0, 1, (n, D))
U[:8192] = 10.0 ** rng.uniform(-300, -1, (8192, D))
U[8192:16384] = 1.0 - 10.0 ** rng.uniform(-16, -1, (8192, D))
U[16384:24576] = 0.5 + rng.uniform(-1e-6, 1e-6, (8192, D))
U = np.clip(U, 1e-300, 1 - 2.0 ** -53)
Ut = torch.from_numpy(U).cuda(); X = torch.empty_like(Ut)
d = model.desc()
_lib.check(_lib.lib().nsb200_transform_batch(ctypes.byref(d), _lib.ptr(Ut), ctypes.c_int64(n), _lib.ptr(X), _lib.stream_arg()))
got, exp = X.cpu().numpy(), ndtri(U)
rel = np.abs(got - exp) / np.maximum(np.abs(exp), 1e-300)
print(f"lib {os.path.basename(_lib._SO)}: max rel err {rel.max():.3e}, mean {rel.mean():.3e}, > 1e-15: {(rel > 1e-15).sum()}")

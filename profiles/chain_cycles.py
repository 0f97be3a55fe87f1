"""Cycle accounting of one chain's critical path (needs the -DNSB_PROFILE build libnsb200_prof.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jaxns_b200 import _lib
_lib._SO = os.path.join(os.path.dirname(_lib._SO), "libnsb200_prof.so")
import jaxns_b200 as j
from jaxns_b200 import distributions as tfpd, likelihoods as lk, random
D = 32
cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)
def prior_model():
    x = yield j.Prior(tfpd.Normal(loc=np.zeros(D), scale=np.ones(D)), name="x")
    return x
m = j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov))
ns = j.NestedSampler(model=m, num_live_points=3200)
L = _lib.lib()
out = (ctypes.c_ulonglong * 16)()
L.nsb200_debug_profile(out, 1)
reason, state = ns(random.PRNGKey(0))
torch.cuda.synchronize()
L.nsb200_debug_profile(out, 0)
v = np.array(list(out), dtype=np.float64)
names = ["prelude", "fetch/direction", "bounds", "proposal gen", "prior transform", "likelihood", "accept/loop", "exit"]
evals, slices = v[8], v[9]
tot = v[:8].sum()
print(f"chain 0 of every iteration: {slices:.0f} slices, {evals:.0f} eval rounds, {tot:.0f} cycles total, {tot/evals:.0f} cycles per eval")
for n, c in zip(names, v[:8]):
    print(f"  {n:18s} {c/tot:6.3f}  {c/evals:8.1f} cyc/eval  {c/slices:8.1f} cyc/slice")
print(f"  evaluations with a lane in the Giles tail (w >= 6.25): {v[10]/evals:.4f}; in the AS241 tail (|q| > 0.425): {v[11]/evals:.4f}")

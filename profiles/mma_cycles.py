"""Cycle accounting of warp 0 of the DMMA slice kernel (needs a -DNSB_PROFILE build: NSB200_LIB=<so> python profiles/mma_cycles.py)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jaxns_b200 import _lib
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
ns(random.PRNGKey(1))
L.nsb200_debug_profile(out, 1)
reason, state = ns(random.PRNGKey(0))
torch.cuda.synchronize()
L.nsb200_debug_profile(out, 0)
v = np.array(list(out), dtype=np.float64)
names = ["prelude", "proposals", "prior transform (central + tail trips)", "DMMA + reduce + logL", "accept / begin_slice"]
rounds = v[8]
tot = v[:5].sum()
it = ns.nested_sampler.last_profile["iterations"]
print(f"warp 0 of every launch: {it} launches, {rounds:.0f} rounds ({rounds/it:.0f} per launch), {tot:.0f} cycles, {tot/rounds:.0f} cycles per round")
for n, c in zip(names, v[:5]):
    print(f"  {n:42s} {c/tot:6.3f}  {c/rounds:8.1f} cyc/round")
print(f"  tail trips per round {v[9]/rounds:.3f}; rounds with an accepting chain {v[10]/rounds:.3f}")

import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import torch, jaxns_b200 as j
from jaxns_b200 import random
from oracle import oracle
from tests.models import product_models, to_oracle
oracle.set_num_threads(os.cpu_count())
name, D, N, S, midpoint = "gauss", 32, 3200, 160, True
model = product_models()[name](D); om = to_oracle(model, oracle)
m = N // 2
for shells in (1, 2, 3):
    sampler = j.UniDimSliceSampler(model=model, num_slices=S, num_phantom_save=0, midpoint_shrink=midpoint, perfect=True)
    ns = j.ShardedStaticNestedSampler(model=model, max_samples=N * 10, init_efficiency_threshold=0.1, sampler=sampler, num_live_points=N)
    reason, register, state = ns._run(random.PRNGKey(42), j.TerminationCondition(max_samples=float(shells * m)))
    ons = oracle.OracleNestedSampler(om, N, S, 0, midpoint, max_samples=N * 10)
    oreason, ost = ons.run(random.PRNGKey(42), oracle.TermCond(max_samples=float(shells * m)))
    n = state.num_samples
    gl = state.sample_collection.log_L[:n].cpu().numpy(); ol = ost["log_L"][:n]
    gn = state.sample_collection.num_likelihood_evaluations[:n].cpu().numpy(); on = ost["n_evals"][:n]
    bad = np.abs(gl - ol) > 1e-6 * np.maximum(1, np.abs(ol))
    print(f"shells={shells}: rows {n}, logL mismatches {bad.sum()}, n_evals mismatches {(gn != on).sum()}, max |dlogL| {np.abs(gl-ol).max():.3e}, first bad rows {np.nonzero(bad)[0][:5]}")

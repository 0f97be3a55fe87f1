import os, sys, time, warnings
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import jaxns_b200 as j
from jaxns_b200 import distributions as tfpd, random
from jaxns_b200.experimental import GlobalOptimisationTerminationCondition
warnings.simplefilter("ignore")
D = 8
def prior_model():
    x = yield j.Prior(tfpd.Uniform(low=-2.0 * np.ones(D), high=2.0 * np.ones(D)), name="x")
    return x
def rosenbrock(x):
    return -(100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2).sum(-1)
model = j.Model(prior_model=prior_model, log_likelihood=rosenbrock)
for g in ("0", "1"):
    os.environ["NSB200_SPLIT_GRAPH_GRAD"] = g
    opt = j.GlobalOptimisation(model=model)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = opt(random.PRNGKey(0), GlobalOptimisationTerminationCondition(max_likelihood_evaluations=2e6, atol=1e-8))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"graph_grad={g}: {dt*1e3:.0f} ms, evals {out.num_likelihood_evaluations}, logL {out.log_L_solution:.3e}, reason {out.termination_reason}")

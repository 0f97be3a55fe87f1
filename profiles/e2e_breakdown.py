"""Wall-clock breakdown of the public-API path on config 2 (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jaxns_b200 as j
from jaxns_b200 import distributions as tfpd, likelihoods as lk, random
from jaxns_b200.internals.tree_structure import SampleTreeGraph, count_crossed_edges
from jaxns_b200.internals.shrinkage_statistics import compute_evidence_stats, logsumexp

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:  # under torchrun: chains sharded over the ranks, as bench.py --gpus N does
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
D = 32
cov = np.full((D, D), 0.99) + 0.01 * np.eye(D)


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


def make():
    def prior_model():
        x = yield j.Prior(tfpd.MultivariateNormalTriL(loc=np.zeros(D), scale_tril=np.eye(D)), name="x")
        return x
    return j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, 15.0), covariance_matrix=cov))


for rep in range(4):
    if world > 1:
        dist.barrier()
    t0 = T(); m = make(); ns = j.NestedSampler(model=m, num_live_points=3200 * world); t1 = T()
    reason, state = ns(random.PRNGKey(rep)); t2 = T()
    sc = state.sample_collection
    n = min(state.num_samples, sc.log_L.numel())
    c = count_crossed_edges(SampleTreeGraph(sc.sender_node_idx[:n], sc.log_L[:n])); t3 = T()
    logL = sc.log_L[:n][c.samples_indices]; U = sc.U_samples[:n][c.samples_indices]; t4 = T()
    fin, per = compute_evidence_stats(logL, c.num_live_points); t5 = T()
    X = m.transform(U); t6 = T()
    lp = m.log_prob_prior(U); t7 = T()
    res = ns.to_results(reason, state); t8 = T()
    host = [res.log_L_samples.cpu(), res.log_dp_mean.cpu(), res.samples["x"].cpu()]; t9 = T()
    print(f"[rank {rank}/{world}] rep {rep}: build {1e3*(t1-t0):.1f} ms | run {1e3*(t2-t1):.1f} | tree {1e3*(t3-t2):.1f} | gather {1e3*(t4-t3):.1f} | "
          f"evidence {1e3*(t5-t4):.1f} | transform {1e3*(t6-t5):.1f} | log_prob_prior {1e3*(t7-t6):.1f} | "
          f"to_results(total) {1e3*(t8-t7):.1f} | d2h {1e3*(t9-t8):.1f} | n={n}")

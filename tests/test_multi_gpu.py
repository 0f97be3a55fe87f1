"""2-GPU test of the sharded loop (needs two visible GPUs; skipped otherwise): the fused NVLink all-gather
(slice kernel stores rows into every rank's gather buffer + device arrival barrier, nsb200_engine_p2p_*), the
host-issued NCCL all-gather and a single-GPU run of the same problem must produce bit-identical dead-point sets
(results are invariant to the number of devices, SURVEY F7 / sharded_static.py:104-129)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import jaxns_b200 as j
    from jaxns_b200 import random
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from models import product_models
    outs = {}
    for name, D, N, k in (("gauss", 8, 512, 0), ("gauss", 32, 4096, 0), ("rosenbrock", 10, 400, 3)):
        model = product_models()[name](D)
        for mode in ("p2p", "nccl", "single"):
            os.environ["NSB200_P2P"] = "0" if mode == "nccl" else "1"
            kw = dict(devices=[0]) if mode == "single" else {}
            ns = j.NestedSampler(model=model, num_live_points=N, max_samples=N * 30, k=k, **kw)
            tc = j.TerminationCondition(max_samples=float(N * 12))
            for rep in range(2):  # the second run re-enters an engine whose peers are already connected
                reason, state = ns(random.PRNGKey(5 + rep), tc)
            if mode == "p2p":
                assert ns.nested_sampler._p2p_state is True, "peer wiring failed"
            sc = state.sample_collection
            n = int(state.num_samples)
            outs[(name, mode)] = (int(reason), n, sc.log_L[:n].clone(), sc.U_samples[:n].clone(),
                                  sc.sender_node_idx[:n].clone(), sc.num_likelihood_evaluations[:n].clone())
            del ns
        for mode in ("nccl", "single"):
            a, b = outs[(name, "p2p")], outs[(name, mode)]
            assert a[0] == b[0] and a[1] == b[1] and a[1] > N * 8, (name, mode, a[:2], b[:2])
            for x, y in zip(a[2:], b[2:]):
                assert torch.equal(x, y), (name, mode)
    # caller-evaluated likelihood (split kernels, host-issued all-gather + NCCL L_min all-reduce every body), plain and
    # with gradient_guided chains (torch-autograd gradients between the kernels): 2 ranks against one
    import warnings
    from jaxns_b200 import distributions as tfpd

    def prior_model():
        x = yield j.Prior(tfpd.Uniform(low=-2.0 * np.ones(4), high=2.0 * np.ones(4)), name="x")
        return x

    def log_likelihood(x):
        return -(100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2).sum(-1)

    model = j.Model(prior_model=prior_model, log_likelihood=log_likelihood)
    for guided in (False, True):
        res = {}
        for mode in ("nccl", "single"):
            kw = dict(devices=[0]) if mode == "single" else {}
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sampler = j.UniDimSliceSampler(model=model, num_slices=8, num_phantom_save=2, midpoint_shrink=True,
                                               perfect=True, gradient_guided=guided)
            ns = j.ShardedStaticNestedSampler(model=model, max_samples=20000, init_efficiency_threshold=0.1,
                                              sampler=sampler, num_live_points=256, **kw)
            reason, reg, state = ns._run(random.PRNGKey(9), j.TerminationCondition(max_samples=4000.0))
            n = int(state.num_samples)
            res[mode] = (int(reason), n, state.sample_collection.log_L[:n].clone(), int(reg.num_likelihood_evaluations))
        a, b = res["nccl"], res["single"]
        assert a[0] == b[0] and a[1] == b[1], (guided, a[:2], b[:2])
        close = torch.isclose(a[2], b[2], rtol=1e-9, atol=1e-9).double().mean().item()
        assert close >= 0.99, (guided, close)
        assert abs(a[3] - b[3]) <= 0.01 * b[3]
    dist.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_fused_all_gather_equals_nccl_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok{r}")) for r in range(2))

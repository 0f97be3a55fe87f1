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

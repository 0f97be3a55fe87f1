"""world_size-2 gloo test (CPU) of the N>1 host path: contiguous chain blocks per rank
(PartitionSpec('shard') of /root/reference/src/jaxns/nested_samplers/sharded/sharded_static.py:104-128)
all-gathered in rank order must reproduce the single-rank batch bit for bit.  The
per-rank compute is done by the oracle here (no GPU); the partition arithmetic, the packed-row
layout and the all-gather are what is under test."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from jaxns_b200.nested_sampler import _dist_info, round_up_num_live_points
    assert _dist_info() == (rank, world)
    D, S, k = 4, 8, 1
    N = round_up_num_live_points(50, 0.5, world)
    m = int(N * 0.5)
    assert m % world == 0
    om = o.gauss_model(D)
    U, logL, _ = o.init_batch(om, o.PRNGKey(3), N)
    order = np.argsort(logL, kind="stable")
    U, logL = U[order], logL[order]
    contour = logL[m - 1]
    key = o.PRNGKey(11)
    per = m // world
    loc = o.slice_batch(om, key, contour, U, logL, S, k, True, num_samples=m, chain_begin=rank * per,
                        chain_end=(rank + 1) * per)
    # packed rows as the engine lays them out: [U[D], logL, nevals, k x (U[D], logL)]
    row = D + 2 + k * (D + 1)
    block = np.zeros((per, row))
    block[:, :D] = loc["U"]
    block[:, D] = loc["log_L"]
    block[:, D + 1] = loc["n_evals"].astype(np.int64).view(np.float64)
    block[:, D + 2:] = np.concatenate([loc["ph_U"].reshape(per, k, D), loc["ph_log_L"].reshape(per, k, 1)],
                                      axis=2).reshape(per, k * (D + 1))
    gathered = torch.zeros((world * per, row), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(block))
    g = gathered.numpy()
    full = o.slice_batch(om, key, contour, U, logL, S, k, True, num_samples=m)
    assert np.array_equal(g[:, :D], full["U"])
    assert np.array_equal(g[:, D], full["log_L"])
    assert np.array_equal(g[:, D + 1].view(np.int64), full["n_evals"])
    assert np.array_equal(g[:, D + 2:].reshape(m, k, D + 1)[:, :, :D].reshape(m * k, D), full["ph_U"])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok{r}")) for r in range(world))


def _agreement_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jaxns_b200.nested_sampler import ContourAgreement
    contour = torch.tensor([-3.25], dtype=torch.float64)
    agree = ContourAgreement(contour, rank)
    for value in (-float("inf"), -3.25, 0.0, 17.5):  # the engine rewrites the scalar in place every body
        contour[0] = value
        agree.all_reduce()
    agree.check()  # identical replicas: silent
    contour[0] = 1.0 + 1e-15 * rank  # one ulp-level drift on rank 1
    agree.all_reduce()
    try:
        agree.check()
        raised = False
    except RuntimeError as e:
        raised = "L_min" in str(e)
    assert raised  # both ranks see that (min, max) differ from their own value or a peer's
    open(os.path.join(out_dir, f"agree{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_contour_agreement_all_reduce(tmp_path):
    """The NCCL L_min agreement of the host-collective path (nested_sampler.ContourAgreement), on gloo."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_agreement_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"agree{r}")) for r in range(world))

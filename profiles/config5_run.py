#!/usr/bin/env python
"""BASELINE config 5 for real: 100-D Gaussian mixture (two isotropic components, benchmarks/difficult_problems style),
num_live_points = 1e5, chains sharded over the GPUs (torchrun), run to dlogZ = log(1 + 1e-3).
  torchrun --nproc-per-node 8 profiles/config5_run.py [num_live] [seed]
Rank 0 prints one JSON line: evals/s, time to termination, iterations, log Z vs the closed form, slice-kernel share."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import jaxns_b200 as j
from jaxns_b200 import random
from tests.models import product_models

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
D = 100
model = product_models()["mixture"](D)
# default max_samples (100 N = 200 shells = 138 nats of compression) would end the run on the sample cap: the narrow
# component needs ~233 nats -- size the store for 500 shells
ns = j.NestedSampler(model=model, num_live_points=N, max_samples=N * 250)
assert ns.num_slices == 500 and ns.k == 0
def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


# engine creation (one ~30 GB arena: 21 GB dead store + 3 x 2.5 GB chain streams) and peer wiring, outside the timed runs
sync()
t0 = time.perf_counter()
inner = ns.nested_sampler
eng = inner._make_engine()
if world > 1:
    inner._connect_peers(eng, world)
sync()
t_engine = time.perf_counter() - t0
runs = []
for rep in range(2):  # run 0 pays lazy module loads, the seed table (36 ms) and the first torch allocations of the state
    sync()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reason, state = ns(random.PRNGKey(seed))
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    runs.append(float(t.item()))
    if rep == 0:
        del state, reason
ms = runs[-1]
prof = ns.nested_sampler.last_profile
reg = ns.nested_sampler.last_register
if rank == 0:
    t1 = time.perf_counter()
    res = ns.to_results(reason, state)
    torch.cuda.synchronize()
    t_res = time.perf_counter() - t1
    # closed form: both components are normalised densities inside the box up to ~3e-4 of their mass
    logZ_true = math.log(2.0) - D * math.log(12.0)
    evals = int(reg.num_likelihood_evaluations)
    rows = ns.nested_sampler.num_live_points // 2
    line = {
        "workload": f"100-D Gaussian mixture, num_live_points={ns.nested_sampler.num_live_points}, num_slices=500, k=0, "
                    f"to dlogZ=log(1+1e-3), {world} GPU(s)",
        "termination_reason": int(reason), "iterations": prof["iterations"], "time_to_termination_ms": ms,
        "likelihood_evals": evals, "evals_per_sec": evals / (ms * 1e-3),
        "slice_kernel_ms_this_rank": prof["slice_ms"], "slice_share": prof["slice_ms"] / ms,
        "ms_per_body": ms / max(1, prof["iterations"]),
        "exchange_bytes_per_body": rows * (D + 2) * 8,
        "log_Z": res.log_Z_mean, "log_Z_uncert": res.log_Z_uncert, "log_Z_closed_form": logZ_true,
        "sigma_off": (res.log_Z_mean - logZ_true) / res.log_Z_uncert, "ESS": res.ESS,
        "total_samples": res.total_num_samples, "to_results_s": t_res, "wall_s": wall, "engine_create_s": t_engine,
        "first_run_ms": runs[0],
    }
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()

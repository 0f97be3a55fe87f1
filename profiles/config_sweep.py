#!/usr/bin/env python
"""Device-timed throughput of the hot path on every BASELINE.json config (the bench line is config 2 only).

For each config: a bounded number of shells (max_samples) or a whole run, CUDA-event timed around
ShardedStaticNestedSampler._run; prints evals/s, ms per iteration, the fused slice kernel's share and its
algorithmic FP64 rate.  usage: python profiles/config_sweep.py [--full] > gpurun_out/config_sweep.txt
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="run configs to termination instead of a shell budget")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    import jaxns_b200 as j
    from jaxns_b200 import random
    from models import product_models
    pm = product_models()
    # name, model, NestedSampler kwargs, flops per eval, shells in the bounded sample
    cfgs = [
        ("cfg1 2-D gauss N=500", pm["gauss"](2), dict(num_live_points=500), 2 * 2 + 4 * 2, None),
        ("cfg2 32-D gauss N=3200", pm["gauss"](32), dict(num_live_points=3200), 32 * 32 + 4 * 32, None),
        ("cfg3 2-D eggbox N=1e4 difficult", pm["eggbox"](2), dict(num_live_points=10000, difficult_model=True), None, None),
        ("cfg4a 10-D rosenbrock difficult+pe", pm["rosenbrock"](10),
         dict(difficult_model=True, parameter_estimation=True), 7 * 9, None),
        ("cfg4b 10-D shells difficult", pm["shells"](10), dict(difficult_model=True), 2 * 30 + 40, None),
        ("cfg5/8 100-D mixture N=12500 (one GPU's share of chains)", pm["mixture"](100),
         dict(num_live_points=12500, max_samples=12500 * 40), 2 * 320, 30),
        ("cfg5 100-D mixture N=1e5 (all chains on one GPU)", pm["mixture"](100),
         dict(num_live_points=100000, max_samples=100000 * 8), 2 * 320, 6),
        ("100-D dense gauss N=3000", pm["gauss"](100), dict(num_live_points=3000, max_samples=3000 * 40),
         100 * 100 + 400, 30),
    ]
    for name, model, kw, flops, shells in cfgs:
        if args.only and args.only not in name:
            continue
        ns = j.NestedSampler(model=model, **kw)
        inner = ns.nested_sampler
        N = inner.num_live_points
        m = int(N * inner.shell_fraction)
        if shells is not None and not args.full:
            tc = j.TerminationCondition(max_samples=float(m * (1 + ns.k) * shells))
        else:
            tc = None
        for rep in range(2):  # rep 0 = warm-up
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            reason, state = ns(random.PRNGKey(rep), tc)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
        n = min(state.num_samples, inner.max_samples)
        evals = int(state.sample_collection.num_likelihood_evaluations[:n].sum().item())
        prof = inner.last_profile
        reg = inner.last_register
        loop_evals = int(reg.num_likelihood_evaluations)
        out = dict(config=name, D=model.U_ndims, N=N, S=ns.num_slices, k=ns.k, iterations=prof["iterations"],
                   ms=ms, ms_per_iter=ms / max(1, prof["iterations"]), evals=evals, evals_per_s=evals / (ms * 1e-3),
                   slice_share=prof["slice_ms"] / ms, slice_evals_per_s=loop_evals / (prof["slice_ms"] * 1e-3),
                   evals_per_slice=loop_evals / max(1, prof["iterations"] * m * ns.num_slices),
                   termination_reason=int(reason))
        if flops:
            out["slice_alg_tflops"] = loop_evals * flops / (prof["slice_ms"] * 1e-3) / 1e12
        if tc is None:
            res = ns.to_results(reason, state)
            out["logZ"] = res.log_Z_mean
            out["logZ_uncert"] = res.log_Z_uncert
        print(json.dumps(out), flush=True)
        del ns, inner, state
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

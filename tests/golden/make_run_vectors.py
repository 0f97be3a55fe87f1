"""Generates tests/golden/run_vectors.json: one tiny whole run of the oracle (4-D correlated Gaussian, 24 live
points, 8 slices, 1 phantom per chain, 6 shells) with its dead-point store, tree counts and evidence results.
The fixture pins the oracle's loop bookkeeping (store offsets, F5 sender quirk, phantom rows, final live append)
against regressions; integers are exact, floats are libm-dependent in the last digits (rtol 1e-9 in the test).
Run: python tests/golden/make_run_vectors.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle as o  # noqa: E402

D, N, S, k = 4, 24, 8, 1
ns = o.OracleNestedSampler(o.gauss_model(D), N, S, k, True, max_samples=N * 40)
reason, st = ns.run(o.PRNGKey(7), max_iterations=6)
r = ns.to_results(reason, st)
n = int(st["num_samples"])
out = dict(
    note="oracle-generated (parity unpinned against live jaxns)", D=D, N=N, S=S, k=k, key=[0, 7], max_iterations=6,
    num_samples=n, next_sample_idx=int(st["next_sample_idx"]), termination_reason=int(reason),
    sender=st["sender"][:n].tolist(), phantom=st["phantom"][:n].astype(int).tolist(),
    n_evals=st["n_evals"][:n].tolist(), log_L=st["log_L"][:n].tolist(),
    U_row0=st["U"][0].tolist(), U_last=st["U"][n - 1].tolist(),
    samples_indices=np.asarray(r["samples_indices"]).tolist(),
    num_live_points_per_sample=np.asarray(r["num_live_points_per_sample"]).tolist(),
    log_Z_mean=float(r["log_Z_mean"]), log_Z_uncert=float(r["log_Z_uncert"]), ESS=float(r["ESS"]),
    H_mean=float(r["H_mean"]), total_num_likelihood_evaluations=int(r["total_num_likelihood_evaluations"]),
    total_phantom_samples=int(r["total_phantom_samples"]))
json.dump(out, open(os.path.join(os.path.dirname(__file__), "run_vectors.json"), "w"), indent=1)
print("wrote run of", n, "samples, reason", reason)

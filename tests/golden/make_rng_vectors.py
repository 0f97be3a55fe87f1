"""Generates tests/golden/rng_vectors.json from the oracle's restatement of jax.random (partitionable
Threefry).  jax itself is not installable in the build container, so these vectors pin the
oracle against regressions and the CUDA path against the oracle; the Random123 KATs and the published
jax words in tests/test_oracle_cpu.py pin the block function itself.
Run: python tests/golden/make_rng_vectors.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle as o  # noqa: E402

cases = []
for seed in (0, 1, 42, 2 ** 32 + 5, 2 ** 63 - 1):
    key = o.PRNGKey(seed)
    cases.append(dict(seed=seed, key=[int(k) for k in key], split4=o.split(key, 4).tolist(),
                      bits64=[int(b) for b in o.random_bits64(key, 4)], uniform=o.uniform(key, 4).tolist(),
                      normal=o.normal(key, 4).tolist()))
json.dump(dict(note="oracle-generated (parity unpinned against live jax)", cases=cases),
          open(os.path.join(os.path.dirname(__file__), "rng_vectors.json"), "w"), indent=1)
print("wrote", len(cases), "cases")

#!/usr/bin/env python
"""Device jax.random.normal against the oracle's (max abs / rel difference), for A/B builds of erfinv's log."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jaxns_b200 import _lib, random
from oracle import oracle as o
n = 1 << 20
key = random.PRNGKey(123)
got = random.normal(key, n).cpu().numpy()
exp = o.normal(key, n)
d = np.abs(got - exp)
print(f"lib {os.path.basename(_lib._SO)}: normals max abs diff {d.max():.3e}, max rel {np.max(d / np.maximum(np.abs(exp), 1e-300)):.3e}, |x|max {np.abs(got).max():.3f}")

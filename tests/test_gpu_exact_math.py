"""Parity build of the slice kernel (-DNSB_EXACT_MATH: Cephes ndtri as TFP evaluates it + IEEE divisions, built by
__graft_entry__.build() as jaxns_b200/libnsb200_exact.so) against the oracle, and the production math on the same
cases: 1.16e6 accept decisions at config-2 size (4 batches of 1600 chains x 160 slices), n_evals equal chain by chain.
The library under test is chosen at import (NSB200_LIB), so each build runs in its own interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch
import jaxns_b200 as j
from jaxns_b200 import _lib, random
from jaxns_b200.types import LivePointCollection
from oracle import oracle
from tests.models import product_models, to_oracle
assert os.path.basename(_lib.so_path()) == %(lib)r, _lib.so_path()
oracle.set_num_threads(os.cpu_count() or 1)
D, N, S = 32, 3200, 160
model = product_models()["gauss"](D)
om = to_oracle(model, oracle)
oU, ologL, _ = oracle.init_batch(om, random.PRNGKey(3), N)
order = np.argsort(ologL, kind="stable")
live_U, live_logL = oU[order], ologL[order]
m = N // 2
sampler = j.UniDimSliceSampler(model=model, num_slices=S, num_phantom_save=0, midpoint_shrink=True, perfect=True)
state = LivePointCollection(None, torch.from_numpy(live_U).cuda(), None, torch.from_numpy(live_logL).cuda(), None)
decisions, worst = 0, 0.0
for seed, rank in ((11, m - 1), (12, m - 1), (13, N // 4), (14, 3 * N // 4)):
    contour = live_logL[rank]
    key = random.PRNGKey(seed)
    exp = oracle.slice_batch(om, key, contour, live_U, live_logL, S, 0, True, num_samples=m)
    sample, _ = sampler.get_samples_batch(key, contour, state, m)
    nev = sample.num_likelihood_evaluations.cpu().numpy()
    assert np.array_equal(nev, exp["n_evals"]), (seed, int((nev != exp["n_evals"]).sum()))
    err = float(np.abs(sample.U_sample.cpu().numpy() - exp["U"]).max())
    assert err < 1e-8, (seed, err)
    worst = max(worst, err)
    decisions += int(nev.sum())
assert decisions > 1000000, decisions
print("OK decisions=%%d max|dU|=%%.2e" %% (decisions, worst))
'''


@pytest.mark.parametrize("lib", ["libnsb200_exact.so", "libnsb200.so"])
def test_million_decisions_match_oracle(lib):
    import torch
    assert torch.cuda.is_available()
    so = os.path.join(ROOT, "jaxns_b200", lib)
    assert os.path.exists(so), f"{so} missing: run __graft_entry__.build()"
    env = dict(os.environ, NSB200_LIB=so)
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, lib=lib)], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "OK decisions=" in out.stdout

#!/usr/bin/env python
"""Final-pass statistics against the HBM roofline (tree count 160 M bytes, evidence stats 80 M bytes):
count_crossed_edges and compute_evidence_stats at M = 2e5 / 1e7, for a dead set in store order (already sorted: a
k = 0 run) and in shuffled order (phantom rows interleaved: the general radix path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jaxns_b200.internals.shrinkage_statistics import compute_evidence_stats
from jaxns_b200.internals.tree_structure import SampleTreeGraph, argsort, count_crossed_edges

PEAK = 6540.5  # GB/s, MEASURED_PEAKS.json


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


for M in (200_000, 2_000_000, 10_000_000):
    g = torch.Generator(device="cuda").manual_seed(M)
    logL = torch.sort(torch.randn(M, dtype=torch.float64, device="cuda", generator=g) * 30 - 100).values
    m = 1600 if M <= 200_000 else 50_000
    # NS-like tree: shell s (m rows) is sent by the second-highest node of shell s - 1
    idx = torch.arange(M, device="cuda")
    sender = torch.clamp((idx // m) * m - 1, min=0)
    for label, perm in (("store order (sorted)", None), ("shuffled", torch.randperm(M, device="cuda", generator=g))):
        if perm is None:
            s2, l2 = sender, logL
        else:
            inv = torch.empty_like(perm)
            inv[perm] = idx
            s2 = torch.where(sender == 0, sender, inv[torch.clamp(sender - 1, min=0)] + 1)[perm].contiguous()
            l2 = logL[perm].contiguous()
        ms, counts = timed(lambda: count_crossed_edges(SampleTreeGraph(s2, l2)))
        ms_sort, _ = timed(lambda: argsort(l2))
        print(f"M={M:9d} {label:22s} count_crossed_edges {ms:8.3f} ms = {160 * M / ms / 1e6:8.1f} GB/s by 160 M "
              f"({160 * M / ms / 1e6 / PEAK:6.1%} of {PEAK}) | argsort alone {ms_sort:8.3f} ms")
    n = counts.num_live_points.to(torch.float64)
    ms, _ = timed(lambda: compute_evidence_stats(logL, n))
    print(f"M={M:9d} compute_evidence_stats (per-sample outputs) {ms:8.3f} ms = {80 * M / ms / 1e6:8.1f} GB/s by 80 M "
          f"({80 * M / ms / 1e6 / PEAK:6.1%})")

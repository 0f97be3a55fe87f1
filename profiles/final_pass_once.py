#!/usr/bin/env python
"""One invocation of the final-pass statistics at M = 1e7 on shuffled input (the general radix path), for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jaxns_b200.internals.shrinkage_statistics import compute_evidence_stats
from jaxns_b200.internals.tree_structure import SampleTreeGraph, count_crossed_edges

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
g = torch.Generator(device="cuda").manual_seed(1)
logL = torch.sort(torch.randn(M, dtype=torch.float64, device="cuda", generator=g) * 30 - 100).values
idx = torch.arange(M, device="cuda")
m = 50_000
sender = torch.clamp((idx // m) * m - 1, min=0)
perm = torch.randperm(M, device="cuda", generator=g)
inv = torch.empty_like(perm)
inv[perm] = idx
s2 = torch.where(sender == 0, sender, inv[torch.clamp(sender - 1, min=0)] + 1)[perm].contiguous()
l2 = logL[perm].contiguous()
torch.cuda.synchronize()
counts = count_crossed_edges(SampleTreeGraph(s2, l2))
final, per = compute_evidence_stats(logL, counts.num_live_points.to(torch.float64))
torch.cuda.synchronize()
print("done", float(final.log_Z_mean))

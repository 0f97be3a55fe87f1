#!/usr/bin/env python
"""Post-process NSB200_TRACE=1 output (stdin): mean slice duration and the mean gap between a slice kernel's end and
the next one's start, with the marks in between."""
import re, sys
import numpy as np
marks = []
for line in sys.stdin:
    m = re.match(r"trace\s+([0-9.]+) us\s+(.*)", line)
    if m:
        marks.append((float(m.group(1)), m.group(2).strip()))
# split into runs: timestamps restart
runs, cur = [], []
for t, name in marks:
    if cur and t < cur[-1][0] - 1000 and name.startswith("merge") is False and t < 2000:
        runs.append(cur); cur = []
    cur.append((t, name))
if cur: runs.append(cur)
for r in runs[-2:]:
    starts = [t for t, n in r if n == "slice start"]
    ends = [t for t, n in r if n == "slice end"]
    k = min(len(starts), len(ends))
    dur = np.array(ends[:k]) - np.array(starts[:k])
    gap = np.array(starts[1:k]) - np.array(ends[:k - 1])
    def after(name):
        ts = [t for t, n in r if n == name]
        out = []
        for e in ends[:k - 1]:
            c = [t for t in ts if t >= e]
            if c: out.append(c[0] - e)
        return np.mean(out[5:]) if len(out) > 5 else float("nan")
    print(f"bodies {k}: slice {dur[5:].mean():.1f} us, gap to next slice {gap[5:].mean():.1f} us | after slice end: merge_rank end +{after('merge_rank end'):.1f}, "
          f"merge_scatter end +{after('merge_scatter end'):.1f}, register update end +{after('register update end'):.1f}, generator end +{after('generator end (side)'):.1f}, "
          f"total {r[-1][0] / 1e3:.2f} ms")

#!/usr/bin/env python
"""Per-CUDA-source-line stall samples and executed instructions from an .ncu-rep captured with
--import-source on.  usage: python profiles/ncu_lines.py rep.ncu-rep [evals_per_launch] [top]"""
import collections
import csv
import subprocess
import sys


def main(path, evals=None, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, res = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] != "" and hdr:
            try:
                ie = hdr.index("Instructions Executed")
                res.append((int(r[4]), int(r[ie]) if r[ie] not in ("", "-") else 0, cur, int(r[0]), r[1].strip()[:80]))
            except (ValueError, IndexError):
                pass
    tot = sum(x[0] for x in res) or 1
    toti = sum(x[1] for x in res)
    print(f"total stall samples {tot}, warp instructions {toti}" + (f" = {toti / evals:.1f} per evaluation" if evals else ""))
    byf = collections.Counter()
    byi = collections.Counter()
    for s, i, f, l, t in res:
        byf[f] += s
        byi[f] += i
    for f in byf:
        print(f"  {f:28s} samples {byf[f] / tot:6.3f}  instructions {byi[f] / toti:6.3f}")
    for s, i, f, l, t in sorted(res, reverse=True)[:top]:
        per = f"{i / evals:7.1f}/eval" if evals else f"{i:10d}"
        print(f"{s / tot:6.3f} {per} {f}:{l}  {t}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 40)

import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
order = []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    if len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    v = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1; agg[name][1] += us
    order.append((name, us))
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.2f} ms over {len(order)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{t/tot:6.3f} {t/1e3:9.3f} ms  n={n:5d}  avg {t/n:9.1f} us  {k[:90]}")

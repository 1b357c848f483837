#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i]
ix = {h: k for k, h in enumerate(hdr)}
t = collections.Counter()
n = collections.Counter()
for r in rows[i + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("s17::", "")
    v = float(r[ix["Metric Value"]])
    u = r[ix["Metric Unit"]]
    v = v / 1e3 if u in ("us", "usecond") else v / 1e6 if u in ("ns", "nsecond") else v * 1e3 if u in ("s", "second") else v
    t[name] += v
    n[name] += 1
tot = sum(t.values())
print("kernel,launches,total_ms,share_pct")
for k, v in t.most_common():
    print("{},{},{:.3f},{:.1f}".format(k, n[k], v, 100 * v / tot))

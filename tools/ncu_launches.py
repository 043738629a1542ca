#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of device time, count, mean."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = r[kn].split("(")[0][:70]
    tot[name] += v
    cnt[name] += 1
S = sum(tot.values())
print(f"launches {sum(cnt.values())}  total {S / 1e6:.3f} ms (ncu: serialised, cold cache -- compare SHARES)")
for k, v in tot.most_common(20):
    print(f"{100 * v / S:6.2f}%  {cnt[k]:6d} x  {v / cnt[k] / 1e3:9.2f} us  {k}")

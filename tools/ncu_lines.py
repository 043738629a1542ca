#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` into per-source-line instruction and
stall-sample shares (first profiled launch only)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot, samp, srcs = collections.Counter(), collections.Counter(), {}
fname, H, cur, seen = None, None, None, set()
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        if fname in seen:       # second launch starts
            break
        seen.add(fname)
        continue
    if "Instructions Executed" in r:
        H = r
        ie, isamp = H.index("Instructions Executed"), H.index("# Samples")
        continue
    if H is None or len(r) <= ie:
        continue
    if r[0] != "":
        cur = (fname, r[0])
        srcs[cur] = r[1]
        continue
    if r[2] in ("...", ""):
        continue
    try:
        tot[cur] += int(r[ie].replace(",", ""))
    except ValueError:
        continue
    try:
        samp[cur] += int(r[isamp].replace(",", ""))
    except ValueError:
        pass
T, S = sum(tot.values()), sum(samp.values())
print("total warp-instructions", T, "stall samples", S)
byfile = collections.Counter()
for (f, l), v in tot.items():
    byfile[f] += v
print({k: round(100 * v / T, 1) for k, v in byfile.most_common(10)})
for k, v in tot.most_common(top):
    print(f"{k[0][:22]:>22}:{k[1]:>5} {100 * v / T:5.1f}% inst {100 * samp[k] / max(S, 1):5.1f}% samp  {srcs[k][:100]}")

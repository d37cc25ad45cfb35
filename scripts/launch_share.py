#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the captured region."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v, u = float(r[vi].replace(",", "")), r[ui]
    v = v / 1e3 if u in ("nsecond", "ns") else v * (1e3 if u in ("msecond", "ms") else 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"{len(rows) - hdr - 1} launches, {s / 1e3:.2f} ms (ncu gpu__time_duration: cold-cache, serialised -> shares, not absolutes)")
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{v / 1e3:10.3f} ms {100 * v / s:5.1f}%  x{cnt[k]:5d}  avg {v / cnt[k]:9.1f} us  {k[:110]}")

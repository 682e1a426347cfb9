#!/usr/bin/env python3
"""Per-kernel share of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  python tools/launch_shares.py <csv>"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; k, v, u = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, n = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= v: continue
    val = float(r[v].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(r[u], 1.0)
    name = r[k].split("(")[0][:64]; tot[name] += val; n[name] += 1
T = sum(tot.values())
for name, t in tot.most_common(14): print(f"{name:64s} {n[name]:5d} launches {t:10.3f} ms {100 * t / T:5.1f} %")

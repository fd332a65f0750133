#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python profiles/launch_summary.py profiles/rNN_ncu_launches_bench.csv > profiles/rNN_ncu_launches_bench_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    us = v / 1e3 if r[mu] in ("ns", "nsecond") else (v if r[mu] in ("us", "usecond") else v * 1e3)
    t = tot.setdefault(r[kn][:60], [0, 0.0])
    t[0] += 1
    t[1] += us
total = sum(t[1] for t in tot.values())
for k, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d total_us=%10.1f share=%5.1f%% avg_us=%8.1f" % (k, cnt, us, 100 * us / total, us / cnt))

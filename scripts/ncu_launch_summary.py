"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: ncu_launch_summary.py <launches.csv> ["comment line"]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[kn].replace("<unnamed>::", "").replace("void ", ""))
    us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}[r[mu]]
    tot[name] += us
    cnt[name] += 1
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# cold-cache, serialised launch times: compare SHARES, not absolutes")
print("kernel,launches,total_us,mean_us,share_pct")
allus = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k},{cnt[k]},{v:.1f},{v / cnt[k]:.1f},{100 * v / allus:.2f}")

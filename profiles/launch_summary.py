"""sum an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name: profiles/launch_summary.py launches.csv [skip_first_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t, n = defaultdict(float), defaultdict(int)
for r in rows[1 + skip:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0][:70]
    t[name] += v
    n[name] += 1
tot = sum(t.values())
for k in sorted(t, key=t.get, reverse=True)[:25]:
    print("%-70s n=%4d  total %9.3f ms  mean %8.3f ms  %5.1f %%" % (k, n[k], t[k], t[k] / n[k], 100 * t[k] / tot))
print("total %.3f ms" % tot)

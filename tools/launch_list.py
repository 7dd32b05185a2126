"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into kernel / launches / total ns / share."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    v = float(r[14].replace(",", ""))
    unit = r[13]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print("kernel, launches, total ns, share")
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-90s %5d %14.1f %6.2f%%" % (name[:90], n, ns, 100 * ns / total))

"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv`
launch list.  Usage: python tools/launch_summary.py launches.csv [header line ...]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
k = next(i for i, r in enumerate(rows) if r[0] == "ID")
h = rows[k]
i_name, i_val, i_unit = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = OrderedDict()
n = 0
for r in rows[k + 1:]:
    if len(r) <= i_val:
        continue
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    a = acc.setdefault(r[i_name], [0.0, 0])
    a[0] += v
    a[1] += 1
    n += 1
total = sum(a[0] for a in acc.values())
for line in sys.argv[2:]:
    print("# " + line)
print(f"# launches {n}, total {total:.1f} us")
for name, (t, c) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:10.1f} us {100 * t / total:5.1f}%  x{c:3d}  avg {t / c:7.1f} us  {name[:70]}")

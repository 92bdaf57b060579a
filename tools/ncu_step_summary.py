"""Per-kernel table from an `ncu --set full` report of a few update steps:
duration, DRAM bytes and achieved DRAM GB/s (against the measured copy peak),
L2 / SM throughput, occupancy.  Kernels launched several times are averaged.
Usage: python tools/ncu_step_summary.py report.ncu-rep [peak_GBps]"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6548.2
if rep.endswith(".csv"):   # `ncu --page raw --csv --log-file x.csv` (metrics-only capture)
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(raw)) if r]
k0 = next(i for i, r in enumerate(rows) if r[0] == "ID")
rows = rows[k0:]
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}


def num(r, name, default=0.0):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0,
                "msecond": 1e3, "second": 1e6}.get(unit, 1.0)


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


acc = OrderedDict()
for r in rows[2:]:
    if len(r) < len(h):
        continue
    name = r[col["Kernel Name"]]
    if "qmcl" not in name:
        continue
    short = name.split("(")[0].replace("void ", "").replace("qmcl::", "").replace("<unnamed>::", "")
    t = to_us(num(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]])
    rd = to_bytes(num(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(num(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
    a = acc.setdefault(short, dict(n=0, t=0.0, rd=0.0, wr=0.0, dram=0.0, l2=0.0, sm=0.0, occ=0.0,
                                   grid=r[col["launch__grid_size"]], block=r[col["launch__block_size"]],
                                   regs=r[col["launch__registers_per_thread"]]))
    a["n"] += 1
    a["t"] += t
    a["rd"] += rd
    a["wr"] += wr
    a["dram"] += num(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    a["l2"] += num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    a["sm"] += num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    a["occ"] += num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")

print(f"# {rep}: per-kernel averages over the captured launches (ncu, cold caches, serialised; under ncu the times are not bench values)")
print(f"# DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration; frac = that / {peak} GB/s "
      "(measured copy peak)")
print(f"{'kernel':44s} {'x':>3s} {'us':>8s} {'grid':>7s} {'blk':>4s} {'regs':>4s} {'DRAM MB':>9s} "
      f"{'GB/s':>7s} {'frac':>5s} {'dram%':>6s} {'L2%':>6s} {'SM%':>6s} {'occ%':>6s}")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1]["t"] / kv[1]["n"]):
    n = a["n"]
    t = a["t"] / n
    by = (a["rd"] + a["wr"]) / n
    gbs = by / (t * 1e-6) / 1e9 if t > 0 else 0.0
    print(f"{k[:44]:44s} {n:3d} {t:8.1f} {a['grid']:>7s} {a['block']:>4s} {a['regs']:>4s} {by / 1e6:9.3f} "
          f"{gbs:7.0f} {gbs / peak:5.2f} {a['dram'] / n:6.1f} {a['l2'] / n:6.1f} {a['sm'] / n:6.1f} "
          f"{a['occ'] / n:6.1f}")

"""Joins an `ncu --page source --csv` export of one kernel (per-SASS-instruction samples and execution
counts) with `nvdisasm -g` line info of the same cubin, and prints instructions / executions / stall
samples per source line.  Usage: sass_lines.py <ncu_source.csv> <nvdisasm_listing> <mangled kernel name>"""
import csv
import re
import sys
from collections import defaultdict

src_csv, listing, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) > ix["stall_no_inst"] and r[0].startswith("0x")]
base = int(data[0][0], 16)
per_off = {}
for r in data:
    def g(n):
        try:
            return int(r[ix[n]] or 0)
        except ValueError:
            return 0
    per_off[int(r[0], 16) - base] = (g("Instructions Executed"), g("# Samples"), g("stall_no_inst"),
                                     g("stall_barrier") + g("stall_sleep"), g("stall_long_sb"))
cur = None
inside = False
agg = defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])  # instrs, executed>0 instrs, warp-execs, samples, no_inst, waiting, long_sb
for line in open(listing, errors="replace"):
    if line.startswith("//--------------------- .text."):
        inside = kernel in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", line)
    if m and cur:
        off = int(m.group(1), 16)
        e = per_off.get(off)
        a = agg[cur]
        a[0] += 1
        if e:
            a[1] += 1 if e[0] else 0
            a[2] += e[0]
            a[3] += e[1]
            a[4] += e[2]
            a[5] += e[3]
            a[6] += e[4]
tot = [sum(a[k] for a in agg.values()) for k in range(7)]
print("total: instrs %d, touched %d, warp-execs %d, samples %d, no_inst %d, waiting %d, long_sb %d" % tuple(tot))
mode = sys.argv[5] if len(sys.argv) > 5 else "touched"
key = {"touched": 1, "instrs": 0, "samples": 3, "no_inst": 4}[mode]
print(f"top {top} source lines by {mode}:")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
    print("%-18s:%-5d instrs %5d touched %5d execs %8d samples %5d no_inst %4d waiting %5d long_sb %4d" % (k[0], k[1], *a))
# per file
pf = defaultdict(lambda: [0] * 7)
for k, a in agg.items():
    for i in range(7):
        pf[k[0]][i] += a[i]
print("per file:")
for k, a in sorted(pf.items(), key=lambda kv: -kv[1][1]):
    print("%-18s instrs %6d touched %6d execs %9d samples %6d no_inst %5d waiting %6d long_sb %5d" % (k, *a))

"""Prints the metrics of an .ncu-rep that the profiles/ summaries quote, and the
warp-stall samples aggregated by code region (regions = runs of SASS lines with
the same execution count).  Usage: python tools/ncu_summary.py report.ncu-rep [kernel-name regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KFILT = ["-k", "regex:" + sys.argv[2]] if len(sys.argv) > 2 else []   # optional kernel-name filter
WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
sm__warps_active.avg.pct_of_peak_sustained_active dram__bytes_read.sum dram__bytes_write.sum
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_lsu_wavefronts.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
smsp__inst_executed.sum sm__issue_active.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
smsp__warps_active.avg.per_cycle_active sm__cycles_elapsed.max smsp__cycles_active.avg""".split()
raw = subprocess.run(["ncu", "-i", rep] + KFILT + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
for name in WANT:
    if name in h:
        i = h.index(name)
        print(f"{name:72s} {units[i]:16s} " + " | ".join(r[i] for r in rows[2:]))
print("# warp stall reasons (per issue-active cycle)")
for i, name in enumerate(h):
    if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio") \
            and "not_issued" not in name:
        v = [float(r[i]) for r in rows[2:]]
        if max(v) >= 0.3:
            print(f"{name:72s} " + " | ".join(f"{x:.2f}" for x in v))
src = subprocess.run(["ncu", "-i", rep] + KFILT + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[k]
R = []
for r in rows[k + 1:]:            # the first captured launch only
    if len(r) < len(h):
        break
    R.append(r)
isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
num = lambda s: int(s or 0)
tot = sum(num(r[isamp]) for r in R)
print(f"# stall samples by code region ({tot} samples, {sum(num(r[iex]) for r in R)} warp instructions)")
ex = [num(r[iex]) for r in R]
segs, start, prev = [], 0, None
for i, e in enumerate(ex):
    if prev is not None and abs(e - prev) > 0.02 * max(e, prev, 1):
        segs.append((start, i, prev))
        start = i
    prev = e
segs.append((start, len(ex), prev))
for a, b, e in segs:
    s = sum(num(r[isamp]) for r in R[a:b])
    if s < 0.005 * tot:
        continue
    st = {h[i][6:]: sum(num(r[i]) for r in R[a:b]) for i in stall}
    top = ", ".join(f"{k} {100 * v / s:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6])
    print(f"  SASS lines {a:4d}-{b:4d} executed {e:7d}x/line: {100 * s / tot:5.1f}% of samples; {top}")

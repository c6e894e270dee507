"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: share of the warp-stall
samples and of the executed instructions.  usage: ncu_lines.py dump.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > ii and r[0].isdigit() and r[si].isdigit() and r[ii].isdigit():
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1][:110]])
        a[0] += int(r[si])
        a[1] += int(r[ii])
ts = sum(a[0] for a in agg.values())
ti = sum(a[1] for a in agg.values())
print("samples", ts, "instructions", ti)
for k in sorted(agg):
    a = agg[k]
    if a[0] > ts * thr / 100 or a[1] > ti * thr / 100:
        print("%-20s %5d  smp %5.1f%%  ins %5.1f%%  %s" % (k[0], k[1], 100 * a[0] / ts, 100 * a[1] / ti, a[2].strip()))

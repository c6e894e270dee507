"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (times are cold-cache and
serialised: the SHARE of a kernel is what carries over to the real step).
usage: launch_summary.py launches.csv [skip_first_n_launches] > profiles/<name>_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
n = 0
for r in rows[1:]:
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[ki])[:90]
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%d launches, %.1f us in total" % (sum(a[0] for a in agg.values()), tot))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.1f %%  %9.1f us  %4d x  %s" % (100 * a[1] / tot, a[1], a[0], name))

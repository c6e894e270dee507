"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass):
UTCHMMA / UTCBAR / LDTM (tcgen05 mma / commit / tensor-memory load), UTMALDG / UBLKCP (TMA tensor / 1-D bulk copies),
HMMA.*TF32 (mma.sync tf32), LDGSTS (cp.async), SYNCS (mbarrier), ATOM / RED.
usage: sass_summary.py path/to/lib.so > profiles/sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

PAT = OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
                   ("HMMA.TF32", r"\bHMMA\.\d+\.F32\.TF32"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("ATOM/RED", r"\b(ATOMG?|RED)\.")])
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, counts = None, OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:100]
        counts[cur] = Counter()
        continue
    if cur:
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
print("%-100s %s" % ("kernel", " ".join("%9s" % k for k in PAT)))
tot = Counter()
for name, c in counts.items():
    if sum(c.values()):
        print("%-100s %s" % (name, " ".join("%9d" % c[k] for k in PAT)))
        tot.update(c)
print("%-100s %s" % ("TOTAL (%d kernels)" % len(counts), " ".join("%9d" % tot[k] for k in PAT)))

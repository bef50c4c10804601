#!/usr/bin/env python
"""Per-SASS-line executed counts of an .ncu-rep source page, grouped by execution count (a loop body
shows up as a run of lines with the same count): how many instructions each loop level costs.
usage: python tools/ncu_loops.py REPORT.ncu-rep [--dump LO HI]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, body = None, []
for r in rows:
    if h is None:
        if "Source" in r and "Instructions Executed" in r:
            h = r
        continue
    if r == h or len(r) < len(h):
        continue
    body.append(r)
si, ex, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(float(r[ex] or 0) for r in body)
tsm = sum(float(r[sm] or 0) for r in body)
print("warp instructions executed %.0f, sass lines %d, samples %.0f" % (tot, len(body), tsm))
c, s = Counter(), Counter()
for r in body:
    c[r[ex]] += 1
    s[r[ex]] += float(r[sm] or 0)
print("exec count x lines -> share of instructions, share of samples")
for k, v in sorted(c.items(), key=lambda kv: -float(kv[0] or 0) * kv[1])[:14]:
    print("%12s x %4d  %5.1f %%  %5.1f %%" % (k, v, 100 * float(k or 0) * v / tot, 100 * s[k] / tsm))
if "--dump" in sys.argv:
    lo, hi = int(sys.argv[sys.argv.index("--dump") + 1]), int(sys.argv[sys.argv.index("--dump") + 2])
    for k in range(lo, min(hi, len(body))):
        r = body[k]
        print("%5d %10s %6s  %s" % (k, r[ex], r[sm], r[si][:120]))

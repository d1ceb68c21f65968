#!/usr/bin/env python
"""Instruction / stall shares of an ncu --set full capture grouped by source-line ranges given as name:file:lo-hi.

    python scripts/phase_shares.py gpurun_out/prof_X.ncu-rep <n_water_frames> name:file:lo-hi ...
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, nwf = sys.argv[1], float(sys.argv[2])
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                      stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(text.splitlines()))
cur, hdr = "", None
agg = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    try:
        inst, samp, th = int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0), int(d.get("Thread Instructions Executed") or 0)
    except ValueError:
        continue
    a = agg[(cur, int(r[0]))]
    a[0] += inst
    a[1] += samp
    a[2] += th
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print("warp-instructions %d (%.1f per water-frame), stall samples %d" % (ti, ti / nwf, ts))
used = set()
for spec in sys.argv[3:]:
    name, fn, rng = spec.split(":")
    lo, hi = (int(x) for x in rng.split("-"))
    keys = [k for k in agg if k[0] == fn and lo <= k[1] <= hi]
    used.update(keys)
    i = sum(agg[k][0] for k in keys)
    s = sum(agg[k][1] for k in keys)
    t = sum(agg[k][2] for k in keys)
    print("%-26s inst %5.1f%%  stall %5.1f%%  threads/inst %4.1f  warp-inst per water-frame %6.1f" % (name, 100 * i / ti, 100 * s / ts, t / max(i, 1), i / nwf))
rest = [k for k in agg if k not in used]
byfile = defaultdict(lambda: [0, 0, 0])
for k in rest:
    for n in range(3):
        byfile[k[0]][n] += agg[k][n]
for fn, (i, s, t) in byfile.items():
    print("%-26s inst %5.1f%%  stall %5.1f%%  threads/inst %4.1f  warp-inst per water-frame %6.1f" % ("(rest) " + fn, 100 * i / ti, 100 * s / ts, t / max(i, 1), i / nwf))

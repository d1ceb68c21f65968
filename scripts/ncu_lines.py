"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
warp-instructions executed and stall samples.  Usage: ncu_lines.py dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = defaultdict(lambda: [0, 0, ""])
cur_file = ""
hdr = None
tot_i = tot_s = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    try:
        inst = int(r[i_inst] or 0)
        samp = int(r[i_samp] or 0)
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    agg[key][0] += inst
    agg[key][1] += samp
    agg[key][2] = r[1].strip()[:90]
    tot_i += inst
    tot_s += samp
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for key, (inst, samp, text) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-18s %4d  inst %5.1f%%  stall %5.1f%%  %s" % (key[0], key[1], 100.0 * inst / max(tot_i, 1), 100.0 * samp / max(tot_s, 1), text))

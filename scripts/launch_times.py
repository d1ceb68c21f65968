"""Print per-kernel device times from an `ncu --metrics gpu__time_duration.sum --csv` log (last N launches)."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = rows[h]
out = []
for r in rows[h + 1:]:
    x = dict(zip(cols, r))
    if x.get("Metric Name") == "gpu__time_duration.sum":
        out.append((int(x["ID"]), x["Kernel Name"].replace("void ", "").replace("wol::", "")[:60], x["Grid Size"], x["Block Size"], float(x["Metric Value"]) / 1e3))
for o in out[-n:]:
    print("%4d %-60s %-14s %-12s %10.1f us" % o)
